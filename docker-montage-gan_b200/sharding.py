"""Batch sharding of the render path across ranks.

Every sample's layer stack and placements are independent, so the path shards on dim 0 with no
data-path collective: rank r of R renders ``x[shard_range(B, r, R)]`` exactly as the reference splits
its batches (``batch_size // num_gpus`` per process, ``custom/training_loop_aio.py:244``).  The only
collectives are bookkeeping: the max-over-ranks step time and the sum of units for the benchmark.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of ``range(total)``: the first ``total % world`` ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def aggregate_throughput(units_this_rank: float, ms_this_rank: float, device=None, group=None):
    """Whole-job throughput = sum over ranks of units / max over ranks of time.  Returns
    (units_total, ms_max, units_per_second).  Works without an initialised process group (N = 1)."""
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([float(ms_this_rank)], dtype=torch.float64, device=device)
        u = torch.tensor([float(units_this_rank)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
        ms, units = float(t.item()), float(u.item())
    else:
        ms, units = float(ms_this_rank), float(units_this_rank)
    return units, ms, units / (ms / 1e3)
