"""Stand-alone multi-GPU driver of the jump engine: the role Kangaroo::Run / SolveKeyGPU (Kangaroo.cpp:913-1083,
510-644) plays in the reference, with the reference's TCP client/server (Network.cpp) replaced in-box by one process
per GPU and an NCCL gather of the DP records to rank 0 (kangaroo_b200/dist.py).  Kept deliberately small: it exists to
exercise the engine end to end (BASELINE config 4) -- the full-featured caller remains the reference's own program
linked against the engine.

  single GPU :  python -m kangaroo_b200.solver in.txt --dp 12
  N GPUs     :  torchrun --nproc-per-node N -m kangaroo_b200.solver in.txt --dp 12
"""
import argparse
import os
import sys
import time

import numpy as np

from . import ecmath as ec
from .engine import GPUEngine, NB_JUMP, NB_RUN, WILD, random_herd_arrays
from .dist import DPGather, decode_records, decode_dp40

ORDER = ec.N


def create_jump_table(range_power):
    """Kangaroo::CreateJumpTable (Kangaroo.cpp:742-832, non-symmetry): 32 distances of rangePower/2+1 bits from MT19937
    seeded 0x600DCAFE (constant "for compatibility of workfiles"), redrawn until the mean is within 2^(jumpBit-1.05 ..
    -0.95); points = d*G.  numpy's RandomState is the same MT19937 with the same seeding, Int::Rand(nbit) consumes
    nbit/32 + 1 words (Int.cpp:988-1001)."""
    jump_bit = min(range_power // 2 + 1, 128)
    rs = np.random.RandomState(0x600DCAFE)
    lo, hi = 2.0 ** (jump_bit - 1.05), 2.0 ** (jump_bit - 0.95)
    nb, left = jump_bit // 32, jump_bit % 32
    for _ in range(100):
        dist = []
        for _i in range(NB_JUMP):
            words = [int(v) for v in rs.randint(0, 2**32, size=nb + 1, dtype=np.uint64)]
            words[nb] &= (1 << left) - 1
            v = sum(w << (32 * k) for k, w in enumerate(words))
            dist.append(v or 1)
        if lo < sum(dist) / NB_JUMP < hi:
            break
    pts = [ec.mul(d) for d in dist]
    return dist, [p[0] for p in pts], [p[1] for p in pts]


def dp_mask(bits):
    """Kangaroo::SetDP (Kangaroo.cpp:154-164)"""
    return 0 if bits == 0 else (~((1 << (64 - min(bits, 64))) - 1)) & 0xFFFFFFFFFFFFFFFF


class Solver:
    def __init__(self, start, end, pub, dp_bits, grid=None, max_found=1 << 17, seed=None, wire="item56"):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.start, self.end, self.pub = start, end, pub
        width = end - start
        self.range_power = width.bit_length()                         # Kangaroo::InitRange (:877-890)
        self.wdiv2 = width >> 1
        # Kangaroo::InitSearchKey (:892-909): keyToSearch = P - start*G
        self.key = ec.add(pub, ec.neg(ec.mul(start))) if start else pub
        gx, gy = grid or GPUEngine.GetGridSize(self.local, 0, 0)
        self.eng = GPUEngine(gx, gy, self.local, max_found)
        self.table = create_jump_table(self.range_power)
        self.eng.SetParams(dp_mask(dp_bits), *self.table)
        self.eng.SetWildOffset(self.wdiv2)
        # Kangaroo::CreateHerd (:670-738): tame d in [0, 2^rangePower), wild d - width/2; points on the device
        n = self.eng.nbKangaroo
        rng = np.random.Generator(np.random.PCG64((seed if seed is not None else int(time.time())) * 1000 + self.rank))
        sc, d128 = random_herd_arrays(n, self.range_power, self.wdiv2, rng)
        self.eng.CreateHerdRaw(sc, d128, self.key)
        self.wire = wire
        self.gather = DPGather(self.eng, self.dist, self.rank, self.world, torch, wire=wire) if self.world > 1 else None
        self.table_dps = {}                                            # rank 0: x -> (d, type)   (HashTable role)
        self.jumps = 0
        self.same_herd = 0

    # rank 0 only -----------------------------------------------------------------------------------------------
    def _check_key(self, td, wd):
        """Kangaroo::CheckKey (Kangaroo.cpp:218-253): the four sign combinations, against key and -key."""
        for t in range(4):
            d1 = (-td) % ORDER if t & 1 else td
            d2 = (-wd) % ORDER if t & 2 else wd
            pk = (d1 + d2) % ORDER
            pt = ec.mul(pk)
            if pt == self.key:
                return (pk + self.start) % ORDER
            if pt == ec.neg(self.key):
                return (-pk + self.start) % ORDER
        return None

    def _insert(self, x, d, ktype):
        """HashTable::Add + Kangaroo::CollisionCheck (Kangaroo.cpp:255-330) on a dict keyed by x."""
        old = self.table_dps.get(x)
        if old is None:
            self.table_dps[x] = (d, ktype)
            return None
        if old[1] == ktype:
            self.same_herd += (old[0] != d)
            return None
        td, wd = (old[0], d) if ktype == WILD else (d, old[0])
        return self._check_key(td, wd)

    def step(self):
        """one Launch on every rank; returns the private key on rank 0 when found (None otherwise)"""
        items = self.eng.Launch()
        self.jumps += self.eng.nbKangaroo * NB_RUN * self.world
        found = None
        if self.gather is not None:
            wo = self.wdiv2
            res = self.gather.step(len(items))
            if self.rank == 0 and self.wire == "dp40":
                # 40-byte records converted on the device (HashTable::Convert): table key = (h, 128 LSBs of x) exactly
                # like the reference's hash table (HashTable.h:52-57), distance already un-biased and signed
                for r, buf in res:
                    for kidx, h, x128, dsigned, ktype in decode_dp40(buf):
                        k = self._insert((h, x128), dsigned % ORDER, ktype)
                        found = found or k
            elif self.rank == 0:
                for r, buf in res:
                    for x, dbias, kidx in decode_records(buf):
                        ktype = kidx % 2
                        d = (dbias - wo) % ORDER if ktype == WILD else dbias       # GPUEngine.cu:672
                        k = self._insert(x, d, ktype)
                        found = found or k
        elif self.rank == 0:
            for it in items:
                k = self._insert(it.x, it.d, it.kIdx % 2)
                found = found or k
        if self.dist is not None:
            flag = self.torch.tensor([1 if found else 0], dtype=self.torch.int32, device="cuda")
            self.dist.broadcast(flag, 0)
            self.stop = bool(flag.item())
        else:
            self.stop = found is not None
        return found

    def run(self, max_steps=1 << 30, verbose=True):
        self.eng.callKernel()
        t0 = time.time()
        key = None
        for s in range(max_steps):
            key = self.step()
            if verbose and self.rank == 0 and (s % 16 == 15 or self.stop):
                dt = time.time() - t0
                print("[%6.1fs] 2^%.2f jumps  %.0f MJump/s  %d DPs  same-herd %d" %
                      (dt, np.log2(max(self.jumps, 1)), self.jumps / dt / 1e6, len(self.table_dps), self.same_herd), flush=True)
            if self.stop:
                break
        self.eng.sync()
        return key


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--dp", type=int, default=12)
    ap.add_argument("--grid", default="")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--max-steps", type=int, default=1 << 30)
    ap.add_argument("--wire", default="item56", choices=["item56", "dp40"], help="DP record format gathered to rank 0")
    a = ap.parse_args(argv)
    start, end, pubs = ec.parse_config(a.config)
    grid = tuple(int(v) for v in a.grid.split(",")) if a.grid else None
    rc = 0
    for i, pub in enumerate(pubs):
        s = Solver(start, end, pub, a.dp, grid, seed=a.seed, wire=a.wire)
        if s.rank == 0:
            print("Range width: 2^%d  kangaroos/GPU: %d  GPUs: %d  dp: %d" % (s.range_power, s.eng.nbKangaroo, s.world, a.dp), flush=True)
        key = s.run(a.max_steps)
        if s.rank == 0:
            if key is not None and ec.mul(key) == pub:
                print("Key#%2d Pub:  0x%064X\n       Priv: 0x%X" % (i, pub[0], key))
            else:
                print("Key#%2d not found" % i); rc = 1
        s.eng.close()
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
