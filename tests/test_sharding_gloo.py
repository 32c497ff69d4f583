"""CPU, world_size 2 over gloo: the multi-GPU path of bench.py shards independent positions across ranks with no data-path
collective; the only exchange is the max-over-ranks reduction of the timed region."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from p3achygo_b200.shard import shard_range, whole_job_rate  # noqa: E402


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8192, 8193):
        for world in (1, 2, 3, 8):
            owned = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                owned.extend(range(lo, hi))
                assert 0 <= hi - lo <= n // world + 1
            assert owned == list(range(n))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    assert whole_job_rate(1024, 10, 8, 0.5) == 1024 * 10 * 8 / 0.5


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import bench
    r, w, local, reduce_max, barrier = bench.dist_setup(world)
    assert (r, w, local) == (rank, world, rank)
    lo, hi = shard_range(8192, r, w)
    barrier()
    slow = reduce_max(1.0 + r)          # max over ranks of a per-rank "time"
    covered = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(covered)            # test-only check that the shards cover the work exactly once
    barrier()
    if r == 0:
        out.put((slow, int(covered.item())))
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_and_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    slow, covered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert slow == 2.0 and covered == 8192
