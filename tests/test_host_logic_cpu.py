"""CPU tests of the stand-alone host logic (kangaroo_b200/ecmath.py, solver.py): no oracle involved in the code under
test; the oracle / reference-generated fixtures are only the checker."""
import numpy as np

from kangaroo_b200 import ecmath as ec
from kangaroo_b200.solver import create_jump_table, dp_mask
from oracle import kgo
from tests.golden_util import load_cases


def test_jump_table_matches_reference_fixtures():
    for c in load_cases():
        dist, px, py = create_jump_table(c["range_power"])
        jd, jpx, jpy = c["table"]
        assert dist == kgo.array_to_ints(jd)
        assert px == kgo.array_to_ints(jpx) and py == kgo.array_to_ints(jpy)


def test_ecmath_vs_oracle(oracle):
    for k in (1, 2, 3, 0xDEADBEEF, ec.N - 1, 2**200 + 12345):
        assert ec.mul(k) == oracle.ec_mul_g(k)
    a, b = ec.mul(11), ec.mul(29)
    assert ec.add(a, b) == ec.mul(40) == oracle.ec_add(a, b)
    assert ec.add(a, ec.neg(a)) is None and ec.add(a, a) == ec.mul(22)


def test_parse_config_and_pubkeys(tmp_path):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    start, end, pubs = ec.parse_config(os.path.join(root, "tests", "golden", "in64.txt"))
    assert end - start == 2**64 - 1
    assert pubs[0] == ec.mul(0x5B3F38AF935A3640D158E871CE6E9666DB862636383386EE510F18CCC3BD72EB)   # README.md:194-195


def test_dp_mask():
    assert dp_mask(0) == 0 and dp_mask(8) == 0xFF00000000000000 and dp_mask(64) == 0xFFFFFFFFFFFFFFFF and dp_mask(99) == 0xFFFFFFFFFFFFFFFF


def test_random_herd_arrays_match_create_herd_rule():
    """Kangaroo::CreateHerd distances (Kangaroo.cpp:696-704), vectorised: tame scalar = v, wild scalar = v - width/2 mod n,
    stored distance = v (wild bias and shift cancel when wildOffset == width/2)."""
    from kangaroo_b200.engine import random_herd_arrays, ORDER
    for rp in (40, 64, 80, 109, 125):
        w = ((1 << rp) - 1) >> 1
        sc, d = random_herd_arrays(2000, rp, w, np.random.Generator(np.random.PCG64(rp)))
        assert sc.shape == (2000, 4) and d.shape == (2000, 2)
        vmax = 0
        for i in range(2000):
            v = int(d[i, 0]) | (int(d[i, 1]) << 64)
            s = sum(int(sc[i, k]) << (64 * k) for k in range(4))
            assert v < (1 << rp)
            assert s == (v if i % 2 == 0 else (v - w) % ORDER)
            vmax = max(vmax, v)
        assert vmax >> (rp - 4)        # the top bits are populated: the range is actually covered
