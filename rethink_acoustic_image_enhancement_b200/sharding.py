"""Batch sharding across the GPUs of one box.

Every image / frame stack / (lq, gt) pair is independent in all three forwards, so the units of a batch are split
into contiguous per-rank slices (64 -> 8 x 8) and processed with **no data-path collective**.  torch.distributed is
used only for rendez-vous, the barrier around timed regions and the max-over-ranks of device timings.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_slice(n_units: int, rank: int, world: int) -> slice:
    """Contiguous slice of `n_units` owned by `rank`; sizes differ by at most one, earlier ranks take the remainder."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(value: float, device: torch.device) -> float:
    """Max of a per-rank scalar (device time of a timed region); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank: int, steps: int, elapsed_ms_this_rank: float, device: torch.device) -> Tuple[float, float]:
    """Whole-job units/s = (sum over ranks of units processed) / (max over ranks of the elapsed time)."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    total = torch.tensor([float(units_per_rank * steps)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    ms = max_over_ranks(elapsed_ms_this_rank, device)
    return float(total.item()) / (ms / 1e3), ms
