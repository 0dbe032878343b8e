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


# ---- host side of the host-buffer step at N > 1: NUMA placement -----------------------------------------------------
# Every rank of `wab_vec_step_host_packed` streams ~1.5 MB per step into pinned host memory. With eight ranks on one
# box the host side decides the scaling: a rank whose threads and pinned pages sit on the other socket pushes all of its
# PCIe traffic over the socket interconnect (round-1 verdict: 0.68 of linear at 8 GPUs with every rank left on node 0).

def parse_cpulist(text: str) -> set:
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the format of sysfs cpulist files)."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_info(pci_bus_id: str, sysfs: str = "/sys") -> Dict[str, object]:
    """NUMA node and local CPUs of a PCI device ('0000:1b:00.0'), from sysfs; node -1 when the platform has none."""
    import os
    base = os.path.join(sysfs, "bus", "pci", "devices", pci_bus_id.lower())
    info: Dict[str, object] = {"pci_bus_id": pci_bus_id.lower(), "numa_node": -1, "local_cpus": set()}
    try:
        with open(os.path.join(base, "numa_node")) as fh:
            info["numa_node"] = int(fh.read().strip())
        with open(os.path.join(base, "local_cpulist")) as fh:
            info["local_cpus"] = parse_cpulist(fh.read())
    except (OSError, ValueError):
        pass
    return info


def _set_mempolicy_preferred(node: int) -> bool:
    """set_mempolicy(MPOL_PREFERRED, {node}): pages this process touches from now on (pinned buffers included) come from
    `node` when it has room. Raw syscall (no libnuma in the image); False when the kernel refuses."""
    import ctypes
    import platform
    nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
    if nr is None or node < 0 or node >= 1024:
        return False
    libc = ctypes.CDLL(None, use_errno=True)
    mask = (ctypes.c_ulong * 16)()                       # 1024 node bits
    mask[node // 64] = 1 << (node % 64)
    MPOL_PREFERRED = 1
    return libc.syscall(ctypes.c_long(nr), ctypes.c_int(MPOL_PREFERRED), mask, ctypes.c_ulong(1024 + 1)) == 0


def bind_to_gpu_numa(device_index: int, pci_bus_id: Optional[str] = None, sysfs: str = "/sys") -> Dict[str, object]:
    """Move this process next to its GPU: CPU affinity = the GPU's local CPUs (those this process is allowed to use),
    memory policy = prefer the GPU's NUMA node. Call it before allocating pinned host buffers. Returns what was done
    (for the benchmark record); a platform without NUMA information is left alone."""
    import os
    if pci_bus_id is None:
        p = torch.cuda.get_device_properties(device_index)
        pci_bus_id = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    info = gpu_numa_info(pci_bus_id, sysfs)
    allowed = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else set()
    local = set(info["local_cpus"]) & set(allowed)
    out = {"pci_bus_id": info["pci_bus_id"], "numa_node": info["numa_node"], "cpus_before": len(allowed),
           "cpus_after": len(allowed), "cpu_bound": False, "mem_bound": False}
    if local and local != set(allowed):
        try:
            os.sched_setaffinity(0, local)
            out["cpu_bound"], out["cpus_after"] = True, len(local)
        except OSError:
            pass
    if int(info["numa_node"]) >= 0:
        out["mem_bound"] = _set_mempolicy_preferred(int(info["numa_node"]))
    return out
