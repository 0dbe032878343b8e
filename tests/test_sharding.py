"""N>1 host logic on CPU: world_size-2 gloo run of the shard arithmetic and the statistics reduce."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from wab_gym_b200.sharding import AsyncStatsReducer, reduce_stats, shard_range


def test_shard_ranges_tile_the_batch():
    for total in (0, 1, 7, 4096, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(1000, rank, world)
    local = torch.tensor([count, 10 * count, rank, 1, 2, 3, 0, first], dtype=torch.int64)
    out[rank] = reduce_stats(local)
    # the asynchronous form used inside timed regions: several submissions in flight, only the last one is read
    red = AsyncStatsReducer(torch.device("cpu"))
    for k in range(1, 4):
        red.submit(lambda k=k: local * k)
    out[100 + rank] = red.result()
    dist.barrier()
    dist.destroy_process_group()


def test_stats_all_reduce_gloo_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == out[1]
    assert out[0]["episodes"] == 1000 and out[0]["steps"] == 10000 and out[0]["finished"] == 1
    assert out[0]["overflows"] == 500 and out[0]["starved"] == 2
    assert out[100] == out[101] and out[100]["episodes"] == 3000 and out[100]["killed"] == 12


def test_async_reducer_without_process_group_returns_local_stats():
    red = AsyncStatsReducer(torch.device("cpu"))
    assert red.result() is None
    red.submit(lambda: torch.arange(8, dtype=torch.int64))
    assert red.result()["steps"] == 1 and red.submitted == 1


def test_numa_binding_follows_the_gpus_sysfs_entry(tmp_path):
    """bind_to_gpu_numa reads the GPU's NUMA node and local CPUs from sysfs, narrows the CPU affinity to the local CPUs
    this process may use and prefers that node's memory; platforms without the information are left alone."""
    import subprocess
    import sys
    from wab_gym_b200.sharding import gpu_numa_info, parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and parse_cpulist("") == set()
    allowed = sorted(os.sched_getaffinity(0))
    dev = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("0\n")
    (dev / "local_cpulist").write_text("%d,100000\n" % allowed[0])
    info = gpu_numa_info("0000:1B:00.0", sysfs=str(tmp_path))
    assert info["numa_node"] == 0 and info["local_cpus"] == {allowed[0], 100000}
    assert gpu_numa_info("0000:ff:00.0", sysfs=str(tmp_path))["numa_node"] == -1
    # in a child process: the affinity change must not leak into the test runner
    code = ("import os, json, sys; sys.path.insert(0, %r); from wab_gym_b200.sharding import bind_to_gpu_numa; "
            "r = bind_to_gpu_numa(0, pci_bus_id='0000:1b:00.0', sysfs=%r); r['now'] = sorted(os.sched_getaffinity(0)); "
            "r2 = bind_to_gpu_numa(0, pci_bus_id='0000:ff:00.0', sysfs=%r); print(json.dumps([r, r2]))"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path), str(tmp_path)))
    import json
    r, r2 = json.loads(subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, check=True, text=True).stdout)
    assert r["numa_node"] == 0 and r["now"] == [allowed[0]]
    assert r["cpu_bound"] == (len(allowed) > 1) and r["cpus_after"] == 1
    assert r2["numa_node"] == -1 and not r2["cpu_bound"] and not r2["mem_bound"]
