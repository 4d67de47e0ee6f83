"""N > 1 host logic on CPU: instance sharding and the result gather over gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hippopt_b200.sharding import gather_instances, gather_rank_rows, shard_range, whole_job_rate


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.tensor([1.0, 10.0], dtype=torch.float64)
        full = gather_instances(local, n)
        expect = torch.arange(n, dtype=torch.float64)[:, None] * torch.tensor([1.0, 10.0], dtype=torch.float64)
        assert torch.equal(full, expect)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        # sharded solves (bench.py::sharded_solves): one row of scalars per rank, whole-job rate = sum / slowest
        row = torch.tensor([100.0 + rank, 90.0 + 5 * rank, 2.0 + rank], dtype=torch.float64)  # instances, converged, seconds
        rows = gather_rank_rows(row)
        assert rows.shape == (world, 3) and torch.equal(rows[rank], row)
        units, slow, rate = whole_job_rate(rows, 1, 2)
        assert (units, slow) == (90.0 + 95.0, 3.0) and rate == pytest.approx(185.0 / 3.0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 8])
def test_gather_over_gloo_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)


def test_rank_rows_without_a_process_group():
    rows = gather_rank_rows(torch.tensor([4.0, 3.0, 0.5], dtype=torch.float64))
    assert rows.shape == (1, 3) and whole_job_rate(rows, 1, 2) == (3.0, 0.5, 6.0)
