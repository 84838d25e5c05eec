"""-m gpu: every compiled kernel variant stays bit-exact: the smoke check (1024 kangaroos x 64 jumps vs the oracle, state and
DP multiset) is run in a subprocess per KGX_MODE / geometry, because the variant is fixed at kgx_create from the environment."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {"KGX_MODE": "stream"},
    {"KGX_MODE": "stream", "KGX_STREAM_G": "64"},
    {"KGX_MODE": "stream", "KGX_STREAM_G": "256"},
    {"KGX_MODE": "resident"},
    {"KGX_MODE": "resident", "KGX_CFG": "64,5"},
    {"KGX_MODE": "resident", "KGX_CFG": "256,3"},
    {"KGX_MODE": "resident", "KGX_CFG": "32,12"},
]


@pytest.mark.parametrize("env", VARIANTS, ids=["-".join(v.values()) for v in VARIANTS])
def test_variant_bit_exact(env):
    e = dict(os.environ); e.update(env)
    p = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "smoke ok" in p.stdout, (p.stdout + p.stderr)[-2000:]
