"""-m gpu: device herd creation (kgx_create_herd) against Kangaroo::CreateHerd as restated by the oracle."""
import numpy as np
import pytest

from kangaroo_b200 import GPUEngine, NB_RUN
from oracle import kgo
from tests.gpu_util import make_case, expected_after

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("range_power", [40, 64, 80, 125])
def test_create_herd_matches_reference_formula(oracle, range_power, kernel):
    n = 2 * 128
    case = make_case(oracle, n, range_power=range_power, seed=1000 + range_power, key_priv=(1 << (range_power - 1)) + 777)
    eng = GPUEngine(2, 1, 0, 65536, **kernel)
    mask = oracle.dp_mask(6)
    eng.SetParams(mask, *case["table"])
    eng.SetWildOffset(case["wdiv2"])
    eng.CreateHerd(kgo.array_to_ints(case["d"]), case["key"])        # same distances, points computed on the GPU
    gx, gy, gd = eng.GetKangaroos()
    assert gx == kgo.array_to_ints(case["px"])
    assert gy == kgo.array_to_ints(case["py"])
    assert gd == kgo.array_to_ints(case["d"])
    eng.callKernel()
    found = eng.Launch(relaunch=False)
    ex, ey, ed, edps = expected_after(oracle, case, NB_RUN, mask)
    gx, gy, gd = eng.GetKangaroos()
    assert gx == kgo.array_to_ints(ex) and gy == kgo.array_to_ints(ey) and gd == kgo.array_to_ints(ed)
    assert sorted((it.x, it.d, it.kIdx) for it in found) == edps
    eng.close()


def test_herd_scalar_edge_cases(oracle, kernel):
    """scalars 1, 2, n-1, powers of two, all-ones pattern: d*G must equal the oracle's k*G."""
    N = kgo.N
    ks = [1, 2, 3, N - 1, N - 2, 2**255 % N, 2**128, 2**64 - 1, 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF] + [2**k for k in range(3, 256, 23)]
    n = 128
    d = [ks[i % len(ks)] for i in range(n)]
    key = oracle.ec_mul_g(0x1234567)
    eng = GPUEngine(1, 1, 0, 1024, **kernel)
    eng.SetWildOffset(0)
    eng.CreateHerd(d, key)
    ax, ay, _ = eng.GetKangaroosRaw()
    for i in range(n):
        exp = oracle.ec_mul_g(d[i]) if i % 2 == 0 else oracle.ec_add(key, oracle.ec_mul_g(d[i]))
        assert (kgo.from_limbs(ax[i]), kgo.from_limbs(ay[i])) == exp, i
    eng.close()


def test_herd_degenerate_terms(oracle, kernel):
    """scalar 0 and the h == 0 cases of the final key addition (d*G == +key -> doubling,
    d*G == -key -> point at infinity) follow the group law instead of reading uninitialised registers."""
    from kangaroo_b200 import ecmath as ec
    kp = 0x1234567
    key = ec.mul(kp)
    n = 128
    d = [(i * 7919 + 11) for i in range(n)]
    d[0] = 0              # tame, scalar 0        -> infinity, stored as (0, 0)
    d[1] = 0              # wild, scalar 0        -> key itself
    d[3] = kp             # wild, d*G == key      -> 2*key
    d[5] = ec.N - kp      # wild, d*G == -key     -> infinity, stored as (0, 0)
    eng = GPUEngine(1, 1, 0, 1024, **kernel)
    eng.SetWildOffset(0)
    eng.CreateHerd(d, key)
    ax, ay, _ = eng.GetKangaroosRaw()
    got = [(kgo.from_limbs(ax[i]), kgo.from_limbs(ay[i])) for i in range(n)]
    assert got[0] == (0, 0)
    assert got[1] == key
    assert got[3] == ec.add(key, key)
    assert got[5] == (0, 0)
    for i in (2, 4, 6, 7, 127):
        exp = ec.mul(d[i]) if i % 2 == 0 else ec.add(key, ec.mul(d[i]))
        assert got[i] == exp, i
    # the engine must keep running from such a herd without faulting
    table = oracle.create_jump_table(64)
    eng.SetParams(oracle.dp_mask(8), *table)
    eng.callKernel()
    eng.Launch(relaunch=False)
    eng.close()
