"""One-process-per-GPU plumbing (torch.distributed).  The stylization path shards by images with no
data-path collective: each rank takes a contiguous slice of the batch; the only
cross-rank operations are the barrier and the max-over-ranks of the device-side timing."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `total` images for `rank`; earlier ranks take the remainder."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """max over all ranks of a per-rank scalar (device-timed milliseconds in bench.py)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
