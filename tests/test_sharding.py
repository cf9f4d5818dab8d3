"""CPU tier: the N > 1 plumbing (scan sharding + whole-job reductions) with world_size = 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcl_augmentation_b200 import sharding


def test_shards_partition_the_stream():
    for n, world in [(4541, 8), (256, 3), (5, 8), (0, 2)]:
        seen = []
        for r in range(world):
            idx = sharding.shard_indices(n, r, world)
            assert all(i % world == r for i in idx)
            seen += idx
        assert sorted(seen) == list(range(n))
    batches = sharding.shard_batches(4541, 3, 8, 256)
    assert sum(len(b) for b in batches) == len(sharding.shard_indices(4541, 3, 8))
    assert max(len(b) for b in batches) <= 256


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_indices(101, rank, world)
        seconds = 1.0 + rank                      # rank 1 is the slow one
        rate = sharding.whole_job_rate(len(mine), seconds)
        total = sharding.all_reduce_scalar(len(mine), "sum")
        dist.barrier()
        out.put((rank, len(mine), rate, total))
    finally:
        dist.destroy_process_group()


def test_whole_job_rate_over_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [51, 50]
    for r in res:
        assert r[3] == 101
        assert r[2] == pytest.approx(101 / 2.0)     # all units / slowest rank's time
