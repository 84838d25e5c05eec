"""Pin the C oracle (oracle/kgx_oracle.c) against the golden vectors of SURVEY.md Appendix C, the known
answers of the reference's sample inputs and Python big-int arithmetic.  CPU only."""
import random

import numpy as np
import pytest

from oracle import kgo

P, N = kgo.P, kgo.N


def test_field_ops_vs_python(oracle):
    rng = random.Random(1)
    for _ in range(2000):
        a, b = rng.randrange(P), rng.randrange(P)
        assert oracle.mod_mul(a, b) % P == a * b % P
        assert oracle.mod_sqr(a) % P == a * a % P
        assert oracle.mod_sub(a, b) == (a - b) % P
    for a in [1, 2, P - 1, 0x1000003D1, 2**255, rng.randrange(P)]:
        assert oracle.mod_inv(a) == pow(a, -1, P)
    assert oracle.mod_inv(0) == 0          # GPUMath.h:785-793 / Int.cpp:1590-1594


def test_fold_has_no_final_subtract(oracle):
    # the fold returns a value < 2^256 congruent mod p, NOT conditionally reduced.
    # a*b = p + 5 exactly happens for a = 1, b = p + 5 (non canonical input, allowed < 2^256).
    assert oracle.mod_mul(1, P + 5) == P + 5
    assert oracle.mod_mul(1, P - 1) == P - 1


def test_order_ops(oracle):
    rng = random.Random(2)
    for _ in range(500):
        a, b = rng.randrange(N), rng.randrange(N)
        assert oracle.order_add(a, b) == (a + b) % N
        assert oracle.order_sub(a, b) == (a - b) % N


def test_mt19937_matches_numpy(oracle):
    # Random.cpp is the classic MT19937 with Knuth seeding == numpy RandomState(seed) raw 32-bit output.
    oracle.rseed(0x600DCAFE)
    ours = [oracle.rndl() for _ in range(1000)]
    rs = np.random.RandomState(0x600DCAFE)
    theirs = [int(v) for v in rs.randint(0, 2**32, size=1000, dtype=np.uint64)]
    assert ours == theirs


def test_rand_bits_consumes_extra_word(oracle):
    # Int.cpp:988-1001: nbit=64 draws THREE words (the third masked to zero)
    oracle.rseed(7)
    w = [oracle.rndl() for _ in range(4)]
    oracle.rseed(7)
    v = oracle.rand_bits(64)
    assert v == w[0] | (w[1] << 32)
    assert oracle.rndl() == w[3]


def test_generator_and_scalar_mult(oracle):
    assert oracle.on_curve(kgo.GX, kgo.GY)
    assert oracle.ec_mul_g(1) == (kgo.GX, kgo.GY)
    x2, y2 = oracle.ec_mul_g(2)
    assert x2 == 0xC6047F9441ED7D6D3045406E95C07CD85C778E4B8CEF3CA7ABAC09B95C709EE5
    assert oracle.ec_mul_g(N - 1) == (kgo.GX, P - kgo.GY)
    # known answer of VC_CUDA8/in64.txt (README.md:194-195): priv -> compressed pubkey 03BB1135...
    priv = 0x5B3F38AF935A3640D158E871CE6E9666DB862636383386EE510F18CCC3BD72EB
    x, y = oracle.ec_mul_g(priv)
    assert x == 0xBB113592002132E6EF387C3AEBC04667670D4CD40B2103C7D0EE4969E9FF56E4 and y & 1 == 1
    # puzzle #110 (puzzle32.txt:6-9)
    x, y = oracle.ec_mul_g(0x35C0D7234DF7DEB0F20CF7062444)
    assert x == 0x09976BA5570966BF889196B7FDF5A0F9A1E9AB340556EC29F8BB60599616167D and y & 1 == 1


def test_jump_table_golden_rangepower64(oracle):
    # SURVEY.md Appendix C
    jd, jpx, jpy = oracle.create_jump_table(64)
    assert oracle.last_draws == 16
    assert kgo.from_limbs(jd[0]) == 0x1D52D864F
    assert kgo.from_limbs(jpx[0]) == 0xE1E1DF42C817E96D60DF6B64D36DEB8798D98AB631348F87A3FF0BD436041502
    assert kgo.from_limbs(jpy[0]) == 0xF705FCC6E23D6F167C48B33E21BF2E53F357A330967774979237228738140769
    assert kgo.from_limbs(jd[1]) == 0x110D4343A
    assert kgo.from_limbs(jpx[1]) == 0xC3F5F6D82FDFB1DF0426FC56FEE9909D9EE3460853388AA3C47B3B49072E505E
    assert kgo.from_limbs(jpy[1]) == 0x2F7D0F7525B738FB27D8BCB4D5B0D9E52C277B9C21A35C10403D4FDDEED77678
    assert kgo.from_limbs(jd[2]) == 0xC00E6F09
    assert kgo.from_limbs(jpx[2]) == 0xEA813C7DAFB6B05D4A46E8DD9B2A7CF5E4F9EF17321212445F64379B5A7FD360
    assert kgo.from_limbs(jpy[2]) == 0x0068A68BB7528985034B8E8B875F30765754045D2B2EA1E7466C715B85374E2E
    assert kgo.from_limbs(jd[31]) == 0xA88B994E


def test_trajectory_golden(oracle):
    # tame kangaroo d0 = 0x123456789ABCDEF, 1024 jumps, dp=8
    table = oracle.create_jump_table(64)
    d0 = 0x123456789ABCDEF
    x0, y0 = oracle.ec_mul_g(d0)
    assert x0 == 0x1A1FD15FCE078234AA292FC024178056BF006433C9B4BD208F59EB4C9EFEC95B
    assert y0 == 0xA18AF1FE46980989D3FF75BF9601121151EF46E2CFAB8999408319CE8F3BE725
    px = kgo.ints_to_array([x0]); py = kgo.ints_to_array([y0]); d = kgo.ints_to_array([d0])
    mask = oracle.dp_mask(8)
    assert mask == 0xFF00000000000000
    dps = oracle.jump_cpu(px, py, d, table, 1, mask, grp=1)
    assert kgo.from_limbs(px[0]) == 0xCBC65663495ABFDEFCE4694628F4981CC143FE64D627E91ECBEB6A40B950960B
    assert kgo.from_limbs(py[0]) == 0xD8B90BB0D444C1CADA427E52086342C4360AE44C7DBE4240BC8033C91731D637
    assert kgo.from_limbs(d[0]) == 0x123456911344336
    dps += [(x, dd, k, j + 1) for (x, dd, k, j) in oracle.jump_cpu(px, py, d, table, 63, mask, grp=1)]
    assert kgo.from_limbs(px[0]) == 0xBE461D05A24484FD1C05981106138A0BEFC199EC07D75A03A45DA8FA9AEA1480
    assert kgo.from_limbs(py[0]) == 0x612666734D842ED168F4DBADEB0CD22D893EF9263999A761A92B1CE0E02F4245
    assert kgo.from_limbs(d[0]) == 0x12345A71FE67C36
    dps += [(x, dd, k, j + 64) for (x, dd, k, j) in oracle.jump_cpu(px, py, d, table, 960, mask, grp=1)]
    assert kgo.from_limbs(px[0]) == 0x00C045C4F68138B38CD06C0F7A75B408012EA52F645DE1A2F9613A486645F4CB
    assert kgo.from_limbs(py[0]) == 0x46D2B91A4BE84A1E84A702B2949178B7A64921B420513256C4CA7B41ED741501
    assert kgo.from_limbs(d[0]) == 0x123496A55C2794D
    assert len(dps) == 7
    by_jump = {j: (x, dd) for (x, dd, k, j) in dps}
    assert by_jump[40] == (0x00F63A24F6B110E87C8B68446ED209DC14A610A73F58B98826FB85174ADC0177, 0x123458DD5A7B79A)
    assert by_jump[204] == (0x0010F6B0D95634CCCD9FF228DCDDF2416E615C4089B5CA92BEC3333583FC89FE, 0x1234635C7CFEC42)
    # invariant d*G == (x, y)
    assert oracle.ec_mul_g(kgo.from_limbs(d[0])) == (kgo.from_limbs(px[0]), kgo.from_limbs(py[0]))


def test_batched_equals_single_and_group_size_independent(oracle):
    # App. A.4: grouping does not change results; Check.cpp:534-549 single-AddDirect formulation agrees.
    table = oracle.create_jump_table(64)
    oracle.rseed(99)
    key = oracle.ec_mul_g(0xDEADBEEFCAFEF00D1234)
    px, py, d = oracle.create_herd(40, 64, 2**63, key)
    ref_states = []
    for i in range(40):
        x, y, dd = kgo.from_limbs(px[i]), kgo.from_limbs(py[i]), kgo.from_limbs(d[i])
        for _ in range(5):
            x, y, dd = oracle.jump_single(x, y, dd, table)
        ref_states.append((x, y, dd))
    a = (px.copy(), py.copy(), d.copy()); b = (px.copy(), py.copy(), d.copy())
    oracle.jump_cpu(*a, table, 5, 0, grp=1024)
    oracle.jump_cpu(*b, table, 5, 0, grp=7)
    for i in range(40):
        assert (kgo.from_limbs(a[0][i]), kgo.from_limbs(a[1][i]), kgo.from_limbs(a[2][i])) == ref_states[i]
    assert all(np.array_equal(u, v) for u, v in zip(a, b))


def test_herd_invariant(oracle):
    # tame: pos = d*G ; wild: pos = key + d*G with d in [-W/2, W/2) mod n  (Kangaroo.cpp:696-728)
    oracle.rseed(5)
    kpriv = 0x1234567
    key = oracle.ec_mul_g(kpriv)
    px, py, d = oracle.create_herd(16, 40, 2**39, key)
    for i in range(16):
        di = kgo.from_limbs(d[i])
        exp = oracle.ec_mul_g(di if i % 2 == 0 else (di + kpriv) % N)
        assert (kgo.from_limbs(px[i]), kgo.from_limbs(py[i])) == exp


def test_gpu_convention_matches_cpu_convention(oracle):
    # device distance = 128-bit, biased by wildOffset on odd kIdx (GPUEngine.cu:407-411, 672)
    table = oracle.create_jump_table(64)
    oracle.rseed(11)
    key = oracle.ec_mul_g(0xABCDEF0123)
    W2 = 2**63
    px, py, d = oracle.create_herd(64, 64, W2, key)
    d128 = np.zeros((64, 2), dtype=np.uint64)
    for i in range(64):
        v = kgo.from_limbs(d[i])
        if i % 2 == 1:
            v = (v + W2) % N
        assert v < 2**128
        d128[i] = kgo.to_limbs(v, 2)
    c = (px.copy(), py.copy(), d.copy())
    g = (px.copy(), py.copy(), d128)
    mask = oracle.dp_mask(4)
    dps_c = oracle.jump_cpu(*c, table, 64, mask)
    dps_g = oracle.jump_gpu_conv(*g, table, 64, mask)
    assert np.array_equal(c[0], g[0]) and np.array_equal(c[1], g[1])
    for i in range(64):
        v = kgo.from_limbs(g[2][i])
        if i % 2 == 1:
            v = (v - W2) % N
        assert v == kgo.from_limbs(c[2][i])
    assert len(dps_c) == len(dps_g) > 100
    unbias = lambda dd, k: (dd - W2) % N if k % 2 else dd
    assert sorted((x, dd, k) for x, dd, k, j in dps_c) == sorted((x, unbias(dd, k), k) for x, dd, k, j in dps_g)


def test_hash_convert(oracle):
    x = 0x1122334455667788_99AABBCCDDEEFF00_0123456789ABCDEF_FEDCBA9876543210
    h, X, D = oracle.hash_convert(x, 5, 1)
    assert h == (x >> 128) & 0x3FFFF and X == x & (2**128 - 1) and D == 5 | (1 << 126)
    h, X, D = oracle.hash_convert(x, N - 5, 0)
    assert D == 5 | (1 << 127)
