"""Pack sharding and the single partial-aggregate exchange of the multi-GPU scan.

Packs are independent units (internal/pack/table/reader.go:299-449 carries no cross-pack state), so
rank r owns the contiguous pack range [r*P/R, (r+1)*P/R) and scans it locally.  The only exchange is
ONE all-gather of a fixed 64-byte partial per aggregate per rank, combined in rank order by
kx_agg_combine (deterministic, float sums stay compensated).  Backend: NCCL on GPUs, gloo in CPU tests.
"""
import ctypes as C

import numpy as np

from .lib import AggOut, lib

PARTIAL_BYTES = C.sizeof(AggOut)   # 48 B payload, exchanged as 64 B


def shard_range(npacks, rank, world):
    """contiguous pack range owned by `rank`"""
    return (rank * npacks) // world, ((rank + 1) * npacks) // world


def pack_partial(agg):
    buf = np.zeros(64, dtype=np.uint8)
    buf[:PARTIAL_BYTES] = np.frombuffer(bytes(agg), dtype=np.uint8)
    return buf


def unpack_partial(buf):
    return AggOut.from_buffer_copy(bytes(bytearray(np.asarray(buf, dtype=np.uint8)[:PARTIAL_BYTES])))


def combine(block_type, parts):
    """fixed-order combine through the C ABI (host-side function, no device needed)"""
    arr = (AggOut * len(parts))(*parts)
    out = AggOut()
    rc = lib().kx_agg_combine(block_type, arr, len(parts), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"kx_agg_combine failed: {rc}")
    return out


def allgather_partials(local_aggs, block_types, dist=None, device=None):
    """local_aggs: list of AggOut of this rank → list of globally combined AggOut.
    One collective for all aggregates of the query."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [combine(t, [a]) for a, t in zip(local_aggs, block_types)]
    import torch
    world = dist.get_world_size()
    n = len(local_aggs)
    mine = np.concatenate([pack_partial(a) for a in local_aggs]) if n else np.zeros(0, dtype=np.uint8)
    t_mine = torch.from_numpy(mine.copy())
    if device is not None:
        t_mine = t_mine.to(device)
    t_all = torch.empty(world * mine.size, dtype=torch.uint8, device=t_mine.device)
    dist.all_gather_into_tensor(t_all, t_mine)
    allb = t_all.cpu().numpy().reshape(world, n, 64)
    return [combine(block_types[j], [unpack_partial(allb[r, j]) for r in range(world)]) for j in range(n)]


def allgather_window_partials(window_aggs, block_types, dist=None, device=None):
    """Sharded series query (kx_scan_buckets per rank): window_aggs[j][k] = this rank's AggOut of value column j in
    window k → the same shape combined over all ranks.  Still ONE collective: the nbuckets x naggs partials travel as
    one flat array and every window cell is combined in rank order like the un-bucketed partials."""
    flat = [a for col in window_aggs for a in col]
    types = [t for col, t in zip(window_aggs, block_types) for _ in col]
    out = allgather_partials(flat, types, dist, device)
    res, i = [], 0
    for col in window_aggs:
        res.append(out[i:i + len(col)])
        i += len(col)
    return res


class PartialExchange:
    """The one collective of a sharded query with every buffer preallocated: the per-rank 64 B partials go
    host (pinned) -> device -> all_gather_into_tensor -> host (pinned) and are combined in rank order."""

    def __init__(self, naggs, dist, device):
        import torch
        self.dist, self.n, self.world = dist, naggs, dist.get_world_size()
        pin = device.type == "cuda"
        self.h_mine = torch.zeros(64 * naggs, dtype=torch.uint8, pin_memory=pin)
        self.h_all = torch.zeros(64 * naggs * self.world, dtype=torch.uint8, pin_memory=pin)
        self.d_mine = torch.zeros(64 * naggs, dtype=torch.uint8, device=device)
        self.d_all = torch.zeros(64 * naggs * self.world, dtype=torch.uint8, device=device)
        self.np_mine = self.h_mine.numpy()
        self.np_all = self.h_all.numpy().reshape(self.world, naggs, 64)

    def exchange(self, local_aggs, block_types):
        for j, a in enumerate(local_aggs):
            self.np_mine[64 * j:64 * j + PARTIAL_BYTES] = np.frombuffer(bytes(a), dtype=np.uint8)
        self.d_mine.copy_(self.h_mine, non_blocking=True)
        self.dist.all_gather_into_tensor(self.d_all, self.d_mine)
        self.h_all.copy_(self.d_all)      # synchronous: the combined result is needed on the host
        return [combine(block_types[j], [unpack_partial(self.np_all[r, j]) for r in range(self.world)]) for j in range(self.n)]
