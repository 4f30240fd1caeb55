"""Host-side helpers for the batch-sharded (replica) multi-GPU run: no data-path collective exists; the
process group is used for start/stop barriers and for the max-over-ranks time."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_batch(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the samples rank `rank` owns when `global_batch` samples are split over `world` ranks."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} does not split over {world} ranks")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank_step: int, steps: int, elapsed_ms_this_rank: float, device=None) -> Tuple[float, float]:
    """Whole-job units/s = (units all ranks processed) / (max over ranks of the elapsed time)."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    ms = max_over_ranks(elapsed_ms_this_rank, device)
    return world * units_per_rank_step * steps / (ms * 1e-3), ms
