"""Data-parallel sharding of an environment batch over the GPUs of one box.

Environments never interact (each reference env owns all of its state, ``wab_env.py:234-238``; v2 worlds are
per-``WAB_Environment2``, ``WAB_Environment2.py:57``), so the batch is cut into contiguous blocks of global env
ids, one per rank, with NO data-path collective. The only exchange is a 64-byte all-reduce of the
episode-statistics vector, issued on a side stream every few launches so that the stream
the step kernels run on never waits for it. Keys of the random draws depend on the GLOBAL env id, so results
are identical for any world size.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

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


def _distributed(group=None) -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def reduce_stats(local_stats: torch.Tensor, group=None) -> Dict[str, int]:
    """Blocking form: sum the int64[8] statistics vector over all ranks (NCCL on GPUs, gloo on CPU) and read it
    on the host. Synchronises the caller's stream — keep it out of any timed or latency-critical region and use
    ``AsyncStatsReducer`` there."""
    import torch.distributed as dist
    t = local_stats.clone()
    if _distributed(group):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_NAMES, (int(v) for v in t.cpu())))


class AsyncStatsReducer:
    """The statistics all-reduce off the critical path.

    ``submit(fetch)`` records an event on the caller's current stream, makes a private side stream wait for it,
    runs ``fetch()`` there (it must enqueue — not synchronise — and return the int64[8] device tensor, e.g.
    ``VecEnv.stats_tensor``) and starts an asynchronous ``all_reduce`` of the result. Nothing is waited for and
    nothing is read on the host until ``result()``; the compute stream is never made to wait. On CPU tensors
    (gloo, the ``-m "not gpu"`` tests) the same calls run without streams.
    """

    def __init__(self, device: Optional[torch.device] = None, group=None):
        self.device = torch.device(device) if device is not None else None
        self.group = group
        self.cuda = self.device is not None and self.device.type == "cuda"
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._pending = None          # (tensor, work or None)
        self.submitted = 0

    def submit(self, fetch: Callable[[], torch.Tensor]) -> None:
        import torch.distributed as dist
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                t = fetch()
                work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if _distributed(self.group) else None
        else:
            t = fetch().clone()
            work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if _distributed(self.group) else None
        self._pending = (t, work)     # an earlier submission still in flight stays ordered before this one on the side stream
        self.submitted += 1

    def result(self) -> Optional[Dict[str, int]]:
        """Totals of the LAST submission (blocks the host until that one collective has finished)."""
        if self._pending is None:
            return None
        t, work = self._pending
        if self.cuda:
            with torch.cuda.stream(self.side):
                if work is not None:
                    work.wait()
                host = t.cpu()
        else:
            if work is not None:
                work.wait()
            host = t
        return dict(zip(STAT_NAMES, (int(v) for v in host)))
