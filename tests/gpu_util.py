"""Shared helpers for the -m gpu parity tests."""
import numpy as np

from oracle import kgo

N = kgo.N


def make_case(oracle, n, range_power=64, seed=1, key_priv=0x5B3F38AF935A3640D158E871CE6E9666DB862636383386EE510F18CCC3BD72EB,
              first_type=0):
    """Jump table + herd exactly as Kangaroo::Check builds them (Check.cpp:472-515) but with a fixed seed."""
    table = oracle.create_jump_table(range_power)
    key = oracle.ec_mul_g(key_priv)
    wdiv2 = ((1 << range_power) - 1) >> 1
    oracle.rseed(seed)
    px, py, d = oracle.create_herd(n, range_power, wdiv2, key, first_type)
    return dict(table=table, key=key, wdiv2=wdiv2, px=px, py=py, d=d, n=n, range_power=range_power)


def cheap_herd(oracle, n, table, seed=3):
    """n distinct valid walkers without n scalar multiplications: start from one random point and take the
    reference jump repeatedly (tame kangaroos: pos = d*G holds by construction)."""
    oracle.rseed(seed)
    d0 = oracle.rand_bits(60) | 1
    x, y = oracle.ec_mul_g(d0)
    px = np.zeros((n, 4), dtype=np.uint64); py = np.zeros((n, 4), dtype=np.uint64); d = np.zeros((n, 4), dtype=np.uint64)
    cx, cy, cd = kgo.ints_to_array([x]), kgo.ints_to_array([y]), kgo.ints_to_array([d0])
    for i in range(n):
        px[i], py[i], d[i] = cx[0], cy[0], cd[0]
        oracle.jump_cpu(cx, cy, cd, table, 1, 0, grp=1)
    return px, py, d


def expected_after(oracle, case, njumps, dp_mask, px=None, py=None, d=None):
    """CPU-convention replay (SolveKeyCPU formula): -> (px, py, d arrays, sorted DP list of (x, d, kidx))"""
    px = (case["px"] if px is None else px).copy()
    py = (case["py"] if py is None else py).copy()
    d = (case["d"] if d is None else d).copy()
    dps = oracle.jump_cpu(px, py, d, case["table"], njumps, dp_mask, grp=1024, max_dp=1 << 22)
    return px, py, d, sorted((x, dd, k) for x, dd, k, j in dps)
