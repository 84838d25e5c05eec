"""The C oracle must reproduce the committed reference-generated fixtures (tests/golden/jump_golden.json)."""
import numpy as np
import pytest

from oracle import kgo
from tests.golden_util import arrays, load_cases

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_golden(oracle, case):
    table = oracle.create_jump_table(case["range_power"])
    assert all(np.array_equal(a, b) for a, b in zip(table, case["table"]))
    oracle.rseed(case["seed"])
    px, py, d = oracle.create_herd(case["n"], case["range_power"], case["width_div2"], case["key"], case["first_type"])
    sx, sy, sd = arrays(case["start"])
    assert np.array_equal(px, sx) and np.array_equal(py, sy) and np.array_equal(d, sd)
    dps = oracle.jump_cpu(px, py, d, table, case["njumps"], case["dp_mask"], grp=128)
    ex, ey, ed = arrays(case["end"])
    assert np.array_equal(px, ex) and np.array_equal(py, ey) and np.array_equal(d, ed)
    assert sorted(dps) == sorted(case["dps"])
    assert len(dps) > 0
