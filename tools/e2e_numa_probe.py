#!/usr/bin/env python
"""Where the host-buffer step's time goes at N GPUs: per-rank device->pinned-host copy bandwidth for the step's 1.5 MB
block, alone (ranks take turns) and with every rank copying at once, with and without NUMA binding; plus the topology.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/e2e_numa_probe.py [--bind]
"""
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    import torch.distributed as dist
    from wab_gym_b200.sharding import bind_to_gpu_numa
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if "--bind" in sys.argv else {"bound": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 4096 * 372
    src = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()

    def copy_rate(iters=300):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            dst.copy_(src, non_blocking=True)
            torch.cuda.current_stream().synchronize()       # one step = one copy + one sync, like the host step
        return nbytes * iters / (time.perf_counter() - t0) / 1e9

    copy_rate(50)
    solo = 0.0
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            solo = copy_rate()
    if world > 1:
        dist.barrier()
    together = copy_rate(600)
    rec = {"rank": rank, "numa": numa, "solo_GBps": round(solo, 2), "together_GBps": round(together, 2),
           "us_per_copy_solo": round(nbytes / solo / 1e3, 2), "us_per_copy_together": round(nbytes / together / 1e3, 2)}
    allrec = [None] * world
    if world > 1:
        dist.all_gather_object(allrec, rec)
    else:
        allrec = [rec]
    if rank == 0:
        for r in allrec:
            print(json.dumps(r))
        if "--topo" in sys.argv:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, text=True).stdout)
            print("allowed cpus:", sorted(os.sched_getaffinity(0)))
            for f in ("/sys/devices/system/node/online", "/sys/devices/system/node/node0/cpulist", "/sys/devices/system/node/node1/cpulist"):
                try:
                    print(f, open(f).read().strip())
                except OSError:
                    pass
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
