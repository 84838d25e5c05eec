"""Multi-GPU distinguished-point gather: independent herds one per GPU, DP records gathered to rank 0.

The reference has no collective: multi-GPU is threads sharing one host hash table under a mutex (Kangaroo.cpp:
594-612) and multi-host is a hand-rolled TCP client/server (Network.cpp).  In-box replacement:
one process per GPU (torch.distributed), no exchange inside the jump loop, and per Launch a variable-length
gather of the 56-byte DP records (GPUMath.h:173-188 layout) to rank 0, which alone owns the hash table:

    counts  : all_gather of one int32 per rank                     (every rank learns max count)
    payload : gather of the first `cap` records of each rank's DP slab, cap = max count rounded up to 256
              records -- equal-sized buffers, so it is a plain NCCL gather over NVLink/NVSwitch (gloo on CPU)

The payload is read straight from the engine's device DP slab (no host bounce); the jump kernel of the NEXT
launch is already running on the engine's own stream while the gather proceeds on torch's stream.
"""
ITEM_BYTES = 56
DP40_BYTES = 40
ROUND = 256


class SlabView:
    """Wraps a raw device pointer as a torch uint8 tensor through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class DPGather:
    def __init__(self, engine, dist, rank, world, torch, device=None, slab_fn=None, max_found=None, wire="item56"):
        """engine: kangaroo_b200.GPUEngine (or None with slab_fn for the CPU/gloo tests).
        slab_fn() -> uint8 tensor [4 + max_found*56] of the most recently completed launch."""
        self.eng, self.dist, self.rank, self.world, self.torch = engine, dist, rank, world, torch
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.max_found = max_found if max_found is not None else engine.maxFound
        self._views = {}
        # wire = "item56": raw engine records; "dp40": the reference's 40-byte DP records, converted on the device
        # (HashTable::Convert) -- 29 % less gather traffic and ready for HashTable::Add(h, x, d)
        self.rec = ITEM_BYTES if wire == "item56" else DP40_BYTES
        self.wire = wire
        self._slab_fn = slab_fn if slab_fn is not None else (self._engine_slab if wire == "item56" else self._engine_dp40)
        self.total_gathered = 0
        self.last = None

    def _engine_slab(self):
        ptr = self.eng.dp_slab_device_ptr()
        v = self._views.get(ptr)
        if v is None:
            nbytes = 4 + self.max_found * ITEM_BYTES
            v = self.torch.as_tensor(SlabView(ptr, nbytes), device=self.device)
            self._views[ptr] = v
        return v

    def _engine_dp40(self):
        ptr = self.eng.convert_dps_device_ptr()
        v = self._views.get(("dp40", ptr))
        if v is None:
            v = self.torch.as_tensor(SlabView(ptr, 4 + self.max_found * DP40_BYTES), device=self.device)
            self._views[("dp40", ptr)] = v
        return v

    def step(self, n_found):
        """Gather this launch's DP records to rank 0.  Returns on rank 0 a list of (rank, uint8 tensor [cnt*56]);
        None elsewhere.  n_found = records this rank produced (capped to max_found here)."""
        torch, dist = self.torch, self.dist
        cnt = min(int(n_found), self.max_found)
        mine = torch.tensor([cnt], dtype=torch.int32, device=self.device)
        counts = torch.empty(self.world, dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(counts, mine)
        counts_h = counts.cpu().tolist()
        cap = max(ROUND, (max(counts_h) + ROUND - 1) // ROUND * ROUND)
        cap = min(cap, self.max_found)
        slab = self._slab_fn()
        local = slab[4:4 + cap * self.rec]
        if self.rank == 0:
            bufs = [torch.empty(cap * self.rec, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
            dist.gather(local, bufs, dst=0)
            out = [(r, bufs[r][:counts_h[r] * self.rec]) for r in range(self.world)]
        else:
            dist.gather(local, None, dst=0)
            out = None
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()   # the slab is recycled by the launch after next
        self.total_gathered += sum(counts_h)
        self.last = out
        return out


def _step_flat(self, n_found):
    """Like step(), but rank 0 receives every rank's records in ONE contiguous device tensor (rank r at offset r*cap*rec),
    so that the whole step's DPs reach the host with a single copy.  Returns (counts list, cap, tensor) on rank 0, None elsewhere."""
    torch, dist = self.torch, self.dist
    cnt = min(int(n_found), self.max_found)
    mine = torch.tensor([cnt], dtype=torch.int32, device=self.device)
    counts = torch.empty(self.world, dtype=torch.int32, device=self.device)
    dist.all_gather_into_tensor(counts, mine)
    counts_h = counts.cpu().tolist()
    cap = max(ROUND, (max(counts_h) + ROUND - 1) // ROUND * ROUND)
    cap = min(cap, self.max_found)
    local = self._slab_fn()[4:4 + cap * self.rec]
    out = None
    if self.rank == 0:
        if getattr(self, "_flat", None) is None:
            self._flat = torch.empty(self.world * self.max_found * self.rec, dtype=torch.uint8, device=self.device)
        n = cap * self.rec
        bufs = [self._flat[r * n:(r + 1) * n] for r in range(self.world)]
        dist.gather(local, bufs, dst=0)
        out = (counts_h, cap, self._flat)
    else:
        dist.gather(local, None, dst=0)
    if self.device.type == "cuda":
        torch.cuda.current_stream().synchronize()   # the slab is recycled by the launch after next
    self.total_gathered += sum(counts_h)
    return out


DPGather.step_flat = _step_flat


def decode_records(buf):
    """uint8 tensor/bytes of 56-byte records -> list of (x, d_biased, kidx) Python ints (GPUEngine.cu:653-671)."""
    b = bytes(buf.cpu().numpy().tobytes()) if hasattr(buf, "cpu") else bytes(buf)
    out = []
    for o in range(0, len(b), ITEM_BYTES):
        x = int.from_bytes(b[o:o + 32], "little")
        d = int.from_bytes(b[o + 32:o + 48], "little")
        k = int.from_bytes(b[o + 48:o + 56], "little")
        out.append((x, d, k))
    return out


def decode_dp40(buf):
    """40-byte DP records -> list of (kIdx32, h, x128, dist (signed int), type): HashTable::CalcDistAndType (HashTable.cpp:249-260)."""
    b = bytes(buf.cpu().numpy().tobytes()) if hasattr(buf, "cpu") else bytes(buf)
    out = []
    for o in range(0, len(b), DP40_BYTES):
        kidx = int.from_bytes(b[o:o + 4], "little")
        h = int.from_bytes(b[o + 4:o + 8], "little")
        x = int.from_bytes(b[o + 8:o + 24], "little")
        d = int.from_bytes(b[o + 24:o + 40], "little")
        ktype = (d >> 126) & 1
        mag = d & ((1 << 126) - 1)
        out.append((kidx, h, x, -mag if d >> 127 else mag, ktype))
    return out
