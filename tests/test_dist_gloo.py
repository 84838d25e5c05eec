"""world_size-2 gloo test (CPU) of the DP gather logic used for the multi-GPU path (kangaroo_b200/dist.py)."""
import os
import socket
import struct

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kangaroo_b200.dist import DPGather, decode_records, ITEM_BYTES


def _records(rank, n):
    out = b""
    for i in range(n):
        x = (rank << 200) | (i * 0x1234567 + 1)
        d = (rank << 100) | i
        out += x.to_bytes(32, "little") + d.to_bytes(16, "little") + struct.pack("<Q", rank * 1000 + i)
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    max_found = 1024
    results = []
    for step, counts in enumerate([(3, 700), (0, 0), (1024, 5), (300, 300)]):
        n = counts[rank]
        slab = torch.zeros(4 + max_found * ITEM_BYTES, dtype=torch.uint8)
        raw = struct.pack("<I", n) + _records(rank, n)
        slab[:len(raw)] = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
        g = DPGather(None, dist, rank, world, torch, device=torch.device("cpu"), slab_fn=lambda s=slab: s, max_found=max_found)
        out = g.step(n)
        if rank == 0:
            got = {r: decode_records(buf) for r, buf in out}
            results.append({r: len(v) for r, v in got.items()})
            for r in range(world):
                exp = decode_records(_records(r, counts[r]))
                assert got[r] == exp, (step, r)
        else:
            assert out is None
    if rank == 0:
        q.put(results)
    dist.destroy_process_group()


def test_dp_gather_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [{0: 3, 1: 700}, {0: 0, 1: 0}, {0: 1024, 1: 5}, {0: 300, 1: 300}]
