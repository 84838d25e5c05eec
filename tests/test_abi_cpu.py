"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/kgx.h declares,
the product fails loudly without a GPU (no CPU fallback), and the host-portable safegcd inverse is correct."""
import ctypes
import os
import random
import re

import pytest

import kangaroo_b200
from kangaroo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "kgx.h")).read()
    declared = set(re.findall(r"\b(kgx_[a-z_0-9]+)\s*\(", hdr))
    lib = kangaroo_b200.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_no_oracle_in_product():
    # the product package must never import / link the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kangaroo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "kgx_oracle" not in src and "libkref" not in src, f


def test_engine_fails_loudly_without_gpu():
    lib = kangaroo_b200.load_library()
    if lib.kgx_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        kangaroo_b200.GPUEngine(2, 32)


def test_host_modinv_safegcd():
    L = ctypes.CDLL(_lib.HOSTTEST_PATH)
    P = 2**256 - 0x1000003D1

    def inv(a):
        i = (ctypes.c_uint64 * 4)(*[(a >> (64 * k)) & (2**64 - 1) for k in range(4)])
        o = (ctypes.c_uint64 * 4)()
        L.kgx_host_modinv(o, i)
        return sum(o[k] << (64 * k) for k in range(4))

    rng = random.Random(1)
    vals = [1, 2, 3, P - 1, P - 2, 2**255, (P + 1) // 2] + [2**k for k in range(256)] + [P - 2**k for k in range(255)]
    vals += [rng.randrange(1, P) for _ in range(5000)]
    for a in vals:
        assert inv(a) == pow(a, -1, P)
    assert inv(0) == 0      # GPUMath.h:785-793


def test_host_fp64_multiplier_is_exact():
    """kgx_field_fp64.cuh (the DFMA-pipe 256x256 -> 512 multiplier, 52-bit limbs, round-toward-zero FMA splitting): the host
    twin must reproduce big-integer products exactly, incl. all-ones and limb-boundary operands."""
    import ctypes
    import random
    from kangaroo_b200 import _lib
    L = ctypes.CDLL(_lib.HOSTTEST_PATH)
    A, O = ctypes.c_uint64 * 4, ctypes.c_uint64 * 8
    M = 2**64 - 1

    def mul(a, b, sq=0):
        o = O()
        L.kgx_host_mul512_fp64(o, A(*[(a >> (64 * i)) & M for i in range(4)]), A(*[(b >> (64 * i)) & M for i in range(4)]), sq)
        return sum(int(o[i]) << (64 * i) for i in range(8))
    random.seed(5)
    edge = [0, 1, 2**256 - 1, 2**255, 2**52 - 1, 2**52, 2**104 - 1, 2**104, 2**208 - 1, (1 << 256) - (1 << 52), 2**256 - 0x1000003D1]
    vals = edge + [random.getrandbits(256) for _ in range(1500)]
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        assert mul(a, b) == a * b
        assert mul(a, a, 1) == a * a
    for a in edge:
        for b in edge:
            assert mul(a, b) == a * b
