"""Data-parallel plumbing of the benchmark / training driver: whole scans are sharded over
ranks (one process per GPU); the only exchange on the path is the gradient all-reduce that
DistributedDataParallel issues (reference: train.py:215-219, pcseg/data/__init__.py:106-113)."""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def scan_seeds(rank: int, pool_index: int, batch: int) -> List[int]:
    """Seeds of the scans rank `rank` processes in pool slot `pool_index`: disjoint across
    ranks and slots (weak scaling: every rank gets `batch` scans of its own)."""
    assert batch <= 20 and pool_index < 50
    return [1000 * rank + 20 * pool_index + i for i in range(batch)]


def max_over_ranks(value_ms: float, device) -> float:
    """Device-timed milliseconds -> max over ranks (a step is as slow as its slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value_ms)
    t = torch.tensor([float(value_ms)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_per_rank: int, world: int, ms: float) -> float:
    """units/s over all ranks given the max-over-ranks time."""
    return units_per_rank * world / (ms / 1e3)
