"""Site-range sharding across the GPUs of one box.

Sites are independent given the (replicated, tiny) pedigree context, so the job is split into
contiguous site ranges, one per rank; every rank produces an ordered shard of results and rank 0
concatenates the shards in rank order.  There is no collective on the data path: the only
communication is the optional final gather of the compact result records / summary counters and the
max-over-ranks reduction of the timing in bench.py.
"""
from __future__ import annotations

import numpy as np


def site_range(n_sites: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) range of rank `rank`; ranges tile [0, n_sites) in rank order."""
    base, extra = divmod(n_sites, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_in_rank_order(local: np.ndarray, dist=None) -> np.ndarray | None:
    """Concatenates per-rank record arrays on rank 0 (all_gather_object keeps this backend agnostic:
    the records are tens of bytes per emitted site, never the packed input)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    if dist.get_rank() != 0:
        return None
    return np.concatenate(parts) if len(parts) else local


def reduce_timing(ms_local: float, sites_local: int, dist=None, device=None):
    """(max over ranks of the device time, total sites) -> whole-job sites/s."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms_local, sites_local, sites_local / (ms_local * 1e-3)
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    n = torch.tensor([sites_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item()), int(n.item()) / (float(t.item()) * 1e-3)
