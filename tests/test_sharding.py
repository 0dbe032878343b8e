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
