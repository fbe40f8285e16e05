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


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import swinwnet_b200 as S
    torch.manual_seed(0)                                   # replicated parameters
    model = torch.nn.ModuleDict({"enc": torch.nn.Linear(6, 5), "dec": torch.nn.Linear(5, 3), "unused": torch.nn.Linear(4, 4),
                                 "frozen": torch.nn.Linear(2, 2)})
    for p in model["frozen"].parameters():
        p.requires_grad_(False)
    red = S.dist.GradReducer(model)
    g = torch.Generator().manual_seed(100 + rank)          # different data shard per rank
    x = torch.randn(7, 6, generator=g)
    h = model["enc"](x)
    loss = (model["dec"](h) ** 2).mean() if rank == 0 else (h ** 2).mean()   # rank 1 never touches "dec": grad None there
    loss.backward()
    # plain lists through the queue: tensors travel as shared-memory handles that die with this process
    local = {n: (None if p.grad is None else p.grad.tolist()) for n, p in model.named_parameters()}
    n_red = red.reduce()
    out.put((rank, n_red, local, {n: (None if p.grad is None else p.grad.tolist()) for n, p in model.named_parameters()}))
    dist.destroy_process_group()


def test_grad_reducer_buckets_average_and_tolerate_missing_grads():
    """config 5 plumbing: per-sub-module buckets, sum / world size, gradients missing on one rank count as zeros,
    gradients missing everywhere stay None, frozen parameters are left alone."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, 29641, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, n0, loc0, red0), (_, n1, loc1, red1) = res
    loc0, red0, loc1, red1 = [{k: (None if v is None else torch.tensor(v)) for k, v in d.items()} for d in (loc0, red0, loc1, red1)]
    assert n0 == n1 == 6 * 5 + 5 + 5 * 3 + 3
    for name in red0:
        if name.startswith(("unused", "frozen")):
            assert red0[name] is None and red1[name] is None
            continue
        a = loc0[name] if loc0[name] is not None else torch.zeros_like(red0[name])
        b = loc1[name] if loc1[name] is not None else torch.zeros_like(red0[name])
        assert torch.allclose(red0[name], (a + b) / 2, atol=1e-7) and torch.equal(red0[name], red1[name])
    assert loc1["dec.weight"] is None and red1["dec.weight"] is not None


def _sync_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import swinwnet_b200 as S
    torch.manual_seed(0)
    model = torch.nn.ModuleDict({"enc": torch.nn.Linear(6, 5), "dec": torch.nn.Linear(5, 3), "ca": torch.nn.Linear(5, 5)})
    sync = S.train.DistributedGradSync(model)             # hooks: every backward ends with the all-reduce
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(100 + rank)
    hist = []
    for step in range(4):                                  # even steps skip "ca" (grad None everywhere), odd steps use it
        x = torch.randn(7, 6, generator=g)
        h = model["enc"](x)
        if step % 2:
            h = h + model["ca"](h)
        loss = (model["dec"](h) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        hist.append((sync.reduced_elements, model["ca"].weight.grad is None))
        opt.step()
    out.put((rank, hist, {n: p.detach().flatten().tolist() for n, p in model.named_parameters()}))
    dist.destroy_process_group()


def test_distributed_grad_sync_hooks_keep_replicas_identical():
    """the even / odd unused-parameter pattern of FullModel_supervised_trainer.py:231-288 through the backward-end hook:
    parameters stay bit-identical on both ranks although every rank sees different data"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, 29651, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, h0, p0), (_, h1, p1) = res
    n_small, n_all = 6 * 5 + 5 + 5 * 3 + 3, 6 * 5 + 5 + 5 * 3 + 3 + 5 * 5 + 5
    assert h0 == h1 == [(n_small, True), (n_all, False), (n_small, True), (n_all, False)]
    for k in p0:
        assert p0[k] == p1[k], k
