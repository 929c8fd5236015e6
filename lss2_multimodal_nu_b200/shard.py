"""Sample-sharded multi-GPU execution of the hot path.

Every term of lift+splat is per sample: the rank's lowest-order digit is the
sample index (reference src/model_baseline.py:106-109), the output's outermost
dimension is the batch (:120) and the path has no learnable parameter, so the
batch is partitioned across ranks with NO data-path collective.  The only
communication is bookkeeping: agreeing on the timed interval (max over ranks)
and, in training, the surrounding model's DDP gradient all-reduce, which this
package does not touch.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of the samples rank `rank` owns; the remainder goes to the low ranks."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_seed(base_seed: int, rank: int, batch_set: int = 0) -> int:
    """Distinct deterministic input seed per (rank, rotating batch set)."""
    return base_seed + 1000 * rank + batch_set


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (the timed interval of a multi-GPU run)."""
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


def gather_floats(value: float, device=None):
    """Every rank's scalar, as a list ordered by rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def aggregate_throughput(samples_this_rank: int, elapsed_ms_this_rank: float, device=None) -> float:
    """Whole-job samples/s: all ranks' samples over the slowest rank's time."""
    total = sum_over_ranks(samples_this_rank, device)
    worst = max_over_ranks(elapsed_ms_this_rank, device)
    return total / (worst * 1e-3)
