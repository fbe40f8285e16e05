"""N>1 host logic on CPU (gloo, world_size 2): batch sharding without a data-path collective, the barrier +
max-over-ranks timing reduction and the whole-job aggregate that bench.py reports."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    lo, hi = bench.shard_bounds(13, rank, world)
    # every rank "processes" its own shard; the job time is the slowest rank's
    my_ms = 10.0 + 5.0 * rank
    dist.barrier()
    job_ms = bench.max_over_ranks(my_ms, dist)
    counts = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([float(hi - lo)]))
    if rank == 0:
        out.put((job_ms, [c.item() for c in counts], bench.whole_job_rate(64, world, 10, job_ms)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29631, q)) for r in range(2)]
    for p in procs:
        p.start()
    job_ms, counts, rate = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert job_ms == 15.0                       # max over ranks, not the mean
    assert counts == [7.0, 6.0] and sum(counts) == 13
    assert abs(rate - 2 * 64 * 10 / 0.015) < 1e-6


def test_shard_bounds_cover_everything():
    import bench
    for n in (1, 7, 64, 4096):
        for world in (1, 2, 4, 8):
            spans = [bench.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
