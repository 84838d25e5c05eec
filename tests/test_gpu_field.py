"""-m gpu: device field arithmetic (kgx_field.cuh, kgx_modinv.h) bit-exact against the oracle through the C ABI."""
import ctypes
import random

import numpy as np
import pytest

import kangaroo_b200
from oracle import kgo

pytestmark = pytest.mark.gpu
P = kgo.P
_u64p = ctypes.POINTER(ctypes.c_uint64)


def run(op, a_vals, b_vals):
    lib = kangaroo_b200.load_library()
    a = kgo.ints_to_array(a_vals); b = kgo.ints_to_array(b_vals)
    out = np.zeros_like(a)
    rc = lib.kgx_test_field(0, op, len(a_vals), a.ctypes.data_as(_u64p), b.ctypes.data_as(_u64p), out.ctypes.data_as(_u64p))
    assert rc == 0, lib.kgx_last_error(None)
    return kgo.array_to_ints(out)


EDGE = [0, 1, 2, 3, P - 1, P - 2, P, P + 1, P + 5, 2**256 - 1, 2**256 - 2, 0x1000003D1, 0x1000003D0, 2**255, 2**128 - 1, 2**128,
        2**32 - 1, 2**32, 2**64 - 1, 2**224 + 1, (P + 1) // 2, 0xFFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000FFFFFFFF00000000]


def pairs():
    rng = random.Random(7)
    a = [x for x in EDGE for _ in EDGE] + [rng.randrange(2**256) for _ in range(20000)] + [rng.randrange(P) for _ in range(20000)]
    b = [y for _ in EDGE for y in EDGE] + [rng.randrange(2**256) for _ in range(20000)] + [rng.randrange(P) for _ in range(20000)]
    return a, b


def test_mul_bit_exact(oracle):
    a, b = pairs()
    got = run(0, a, b)
    for x, y, g in zip(a, b, got):
        assert g == oracle.mod_mul(x, y), (hex(x), hex(y))


def test_sqr_bit_exact(oracle):
    a, _ = pairs()
    got = run(1, a, a)
    for x, g in zip(a, got):
        assert g == oracle.mod_sqr(x), hex(x)


def test_sub_bit_exact(oracle):
    a, b = pairs()
    got = run(2, a, b)
    for x, y, g in zip(a, b, got):
        assert g == oracle.mod_sub(x, y), (hex(x), hex(y))


def test_inv_canonical(oracle):
    rng = random.Random(9)
    a = [0, 1, 2, P - 1, P - 2, 2**255, (P + 1) // 2] + [2**k for k in range(0, 256, 3)] + [rng.randrange(1, P) for _ in range(4000)]
    got = run(3, a, a)
    for x, g in zip(a, got):
        assert g == (pow(x, -1, P) if x else 0), hex(x)
