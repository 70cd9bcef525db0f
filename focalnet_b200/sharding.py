"""Batch-axis sharding of the SS2D path across ranks (one process per GPU).

The path has no cross-sample coupling: every (batch, direction, channel) recurrence is independent, so
ranks own contiguous slices of the batch axis and no data-path collective is needed.  The only collectives are the
ones of the harness: a barrier around the timed region and a MAX over ranks of the device time."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of `n_items` units (images / tiles) owned by `rank`; sizes differ by at most 1."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank: float, seconds_this_rank: float, device="cpu") -> float:
    """Whole-job rate = units all ranks processed / the slowest rank's time."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    total = units_per_rank
    if world > 1:
        t = torch.tensor([float(units_per_rank)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total = float(t.item())
    return total / max_over_ranks(seconds_this_rank, device)
