"""Data-parallel sharding of an environment batch over the GPUs of one box.

Environments never interact (each reference env owns all of its state, ``wab_env.py:234-238``), so
the batch is cut into contiguous blocks of global env ids, one per rank, with NO data-path
collective. The only exchange is a 64-byte all-reduce of the episode-statistics vector.
Keys of the random draws depend on the GLOBAL env id, so results are identical for any world size.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .vec_env import STAT_NAMES


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first global env id, count) of `rank`'s contiguous shard; the first `total % world` ranks get one more."""
    if not (0 <= rank < world_size) or total_envs < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total_envs, world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def reduce_stats(local_stats: torch.Tensor, group=None) -> Dict[str, int]:
    """Sum the int64[8] statistics vector over all ranks (NCCL on GPUs, gloo on CPU). Returns a dict."""
    import torch.distributed as dist
    t = local_stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_NAMES, (int(v) for v in t.cpu())))
