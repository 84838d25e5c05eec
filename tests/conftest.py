import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def oracle():
    from oracle import kgo
    return kgo.Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle import kgo
    if not kgo.reference_available():
        pytest.skip("reference build (oracle/_ref/libkref.so) not available")
    return kgo.Reference()
