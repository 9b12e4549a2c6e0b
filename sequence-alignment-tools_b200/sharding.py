"""Partitioning a database stream over ranks (one process per GPU) and gathering
the per-rank hit lists.  No data-path collective: every rank scans its own
shard; torch.distributed (NCCL on the GPUs, gloo in the CPU tests) only moves
hit counts and the compacted hit records at the end.

Ownership rule (include/pm_b200.h, pm_find_patterns): a hit belongs to the rank
whose [begin, end) contains the FIRST raw candidate position of its chain.
A shard therefore carries `overlap` = longest pattern + 3k + 2 characters of
left context (candidate warm-up m+k, chain look-back 2k+1, alignment reach) and
`right_guard` characters of right context so that a chain starting just before
`end` can be followed to its last candidate.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    begin: int     # owned candidate positions [begin, end)
    end: int
    load_lo: int   # stream[load_lo:load_hi] is what the rank loads (offset = load_lo)
    load_hi: int


def shard_plan(stream_len: int, world: int, overlap: int, right_guard: int = 1 << 16,
               cut_points=None):
    """Even split of candidate positions 0..stream_len (inclusive) over `world`
    ranks.  `cut_points` (sorted stream positions, e.g. sequence separators) makes
    the cuts snap to the nearest sequence boundary at or after the even cut."""
    bounds = [0]
    for r in range(1, world):
        b = (stream_len + 1) * r // world
        if cut_points is not None and len(cut_points):
            i = int(np.searchsorted(cut_points, b, side="left"))
            if i < len(cut_points):
                b = int(cut_points[i])
        bounds.append(max(b, bounds[-1]))
    bounds.append(stream_len + 1)
    shards = []
    for r in range(world):
        begin, end = bounds[r], bounds[r + 1]
        lo = max(0, begin - overlap) if r else 0
        hi = min(stream_len, end + right_guard)
        shards.append(Shard(r, begin, end, lo, hi))
    return shards


def gather_hits(local: np.ndarray, dist=None):
    """All ranks contribute a structured hit array; every rank gets the
    concatenation sorted by (first, pattern_id).  Two collectives: counts, then the
    records as bytes (padded to the largest count)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world = dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    cnt = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    item = local.dtype.itemsize
    mx = max(counts)
    if mx == 0:
        return local
    buf = torch.zeros(mx * item, dtype=torch.uint8, device=dev)
    if len(local):
        buf[:len(local) * item] = torch.from_numpy(np.frombuffer(local.tobytes(), np.uint8).copy()).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    parts = [np.frombuffer(b.cpu().numpy().tobytes()[:c * item], dtype=local.dtype) for b, c in zip(bufs, counts)]
    out = np.concatenate(parts)
    return out[np.lexsort((out["pattern_id"], out["first"]))]


class _DeviceBytes:
    """A device pointer as a CUDA-array-interface object (so that torch can wrap it)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def gather_hits_device(d_ptr: int, n_hits: int, item_bytes: int, dist, device=None):
    """The hit lists of all ranks, as they lie on the devices (pm_find_patterns_device),
    gathered over NCCL without a detour through the host.  Returns (counts, records): a
    uint8 device tensor holding the records of rank 0, rank 1, ... back to back.  The
    shards of shard_plan() own disjoint, ascending ranges of `first`, and every rank's list
    is sorted by (first, pattern_id): the concatenation in rank order is the sorted list of
    the whole database.  Two collectives: the counts, then the records padded to the largest
    count (all_gather_into_tensor)."""
    import torch
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    nbytes = n_hits * item_bytes
    local = torch.as_tensor(_DeviceBytes(d_ptr, nbytes), device=dev) if nbytes else torch.empty(0, dtype=torch.uint8, device=dev)
    if world == 1:
        return [n_hits], local
    cnt = torch.tensor([n_hits], dtype=torch.int64, device=dev)
    counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, cnt)
    counts = [int(x) for x in counts_t.tolist()]
    mx = max(counts)
    if mx == 0:
        return counts, torch.empty(0, dtype=torch.uint8, device=dev)
    send = torch.zeros(mx * item_bytes, dtype=torch.uint8, device=dev)
    send[:nbytes] = local
    recv = torch.empty(world * mx * item_bytes, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send)
    parts = [recv[r * mx * item_bytes: r * mx * item_bytes + counts[r] * item_bytes] for r in range(world)]
    return counts, torch.cat(parts)
