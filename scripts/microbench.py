"""Raw device microbenchmarks (BASELINE config 5 inputs): IMAD.WIDE issue rate, fe_mul / fe_sqr / inverse
throughput, and jump-kernel launch time on a synthetic herd.  Writes gpurun_out/microbench.json."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kangaroo_b200  # noqa: E402
from kangaroo_b200 import GPUEngine, NB_RUN  # noqa: E402


def raw(kind, iters):
    lib = kangaroo_b200.load_library()
    ms, ops = ctypes.c_float(0), ctypes.c_double(0)
    assert lib.kgx_bench_raw(0, kind, iters, ctypes.byref(ms), ctypes.byref(ops)) == 0
    return ms.value, ops.value


def main():
    out = {}
    for name, kind, iters in (("imad_wide", 0, 20000), ("fe_mul", 1, 2000), ("fe_sqr", 2, 2000), ("fe_inv", 3, 20)):
        best = None
        for _ in range(3):
            ms, ops = raw(kind, iters)
            r = ops / (ms * 1e-3)
            best = r if best is None or r > best else best
        out[name + "_per_s"] = best
        print(name, "%.3e /s" % best, flush=True)
    # jump kernel on a synthetic herd: every kangaroo = a valid point; use replicated walkers from the oracle-free path:
    # G multiples are not available without the oracle, so use points generated on device by prior jumps: start all at G.
    GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
    GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
    sys.path.insert(0, ROOT)
    from oracle import kgo
    orc = kgo.Oracle()
    table = orc.create_jump_table(80)
    for grid in ((296, 128), (74, 128), (16, 128)):
        n = grid[0] * grid[1] * 128
        eng = GPUEngine(grid[0], grid[1], 0, 1 << 20)
        eng.SetParams(orc.dp_mask(16), *table)
        base = 1024
        from tests.gpu_util import cheap_herd
        bx, by, bd = cheap_herd(orc, base, table, seed=5)
        ax = np.tile(bx, (n // base, 1)); ay = np.tile(by, (n // base, 1)); ad = np.tile(bd[:, :2], (n // base, 1))
        eng.SetKangaroosRaw(ax, ay, ad)
        eng.callKernel()
        times = []
        for i in range(6):
            eng.Launch()
            times.append(eng.last_launch_ms())
        eng.sync()
        ms = min(times[1:])
        out["jump_%dx%d" % grid] = {"kangaroos": n, "ms_per_launch": ms, "mjump_s": n * NB_RUN / ms / 1e3, "all_ms": times}
        print(grid, out["jump_%dx%d" % grid], flush=True)
        eng.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
