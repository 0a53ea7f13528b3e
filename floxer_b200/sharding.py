"""Read sharding for N processes (one per GPU): the path has no exchange step, so a rank verifies
a contiguous block of reads against its own replica of the packed reference, and rank 0 gathers the per-read
records in input order.  torch.distributed is plumbing only (gather of Python objects, reduction of timings)."""
from __future__ import annotations

import numpy as np

from .batch import ReadBatch


def shard_bounds(batch: ReadBatch, world_size: int) -> list[tuple[int, int]]:
    """Contiguous read ranges with similar numbers of anchors (verification work is per anchor)."""
    work = (batch.reads["num_anchors_forward"].astype(np.int64) + batch.reads["num_anchors_reverse"]) + 1
    csum = np.concatenate([[0], np.cumsum(work)])
    cuts = [int(np.searchsorted(csum, csum[-1] * r / world_size, side="left")) for r in range(world_size)] + [len(batch)]
    cuts = np.maximum.accumulate(np.minimum(cuts, len(batch)))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world_size)]


def shard(batch: ReadBatch, rank: int, world_size: int) -> tuple[ReadBatch, int]:
    """This rank's reads and the index of its first read in the whole batch."""
    lo, hi = shard_bounds(batch, world_size)[rank]
    return batch.slice(lo, hi), lo


def gather_records(records, first_read: int, dist=None, dst: int = 0):
    """records = [(local_read_index, ...)] of this rank -> on rank dst the records of all ranks with global read
    indices, in read order; None elsewhere.  dist = torch.distributed (already initialised) or None for one process."""
    mine = [(r[0] + first_read,) + tuple(r[1:]) for r in records]
    if dist is None or dist.get_world_size() == 1:
        return mine
    gathered = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(mine, gathered, dst=dst)
    if dist.get_rank() != dst:
        return None
    return [rec for part in gathered for rec in part]


def reduce_timing(ms: float, units: float, dist=None, device=None):
    """(max over ranks of ms, sum over ranks of units) -- the contract's whole-job throughput ingredients."""
    if dist is None or dist.get_world_size() == 1:
        return ms, units
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
