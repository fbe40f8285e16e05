"""BASELINE configs[4]: the reference's FullModelTrainer (UNMODIFIED, baseline/_ref) driving the drop-in SwinWNet for even+odd
step pairs (FullModel_supervised_trainer.py:199-211, 231-288), data-parallel over the ranks of one box:

  forward   sm_100a kernels (16-bit tensor-core operands, fp32 accumulation) under the trainer's autocast + GradScaler
  backward  KernelOp recompute through ATen (autograd.py) — hand-written backward kernels are not built yet
  exchange  ONE NCCL all-reduce of the flat fp32 gradient bucket per backward (train.DistributedGradSync)
  update    ONE fused AdamW launch (train.FusedAdamW)

    python tools/train_step_bench.py [--batch 8] [--pairs 3]            # 1 GPU
    torchrun --nproc-per-node 8 tools/train_step_bench.py               # 8 GPUs

Rank 0 prints one JSON line: training diffractions/s over all ranks (max over ranks, CUDA events + barrier), the losses
of the first even / odd step next to the reference model's own losses on the same batch and weights, the all-reduced
element count per backward, and a replica-consistency check (parameter checksum equal on all ranks after training).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import benchdata  # noqa: E402
from stage_reference import import_reference  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8, help="diffractions per GPU per step")
    ap.add_argument("--pairs", type=int, default=3, help="timed even+odd step pairs")
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import swinwnet_b200 as S
    _, R, _ = import_reference()
    from FullModel_supervised_trainer import FullModelTrainer

    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    sd = benchdata.make_state_dict(man["wnet_em"], seed=1)
    n_steps = 2 * (a.warmup + a.pairs)
    x = benchdata.synthetic_diffractions(min(a.batch, 4), seed=300 + rank, two_channel=False)
    x = x.repeat((a.batch + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:a.batch]
    xs = torch.cat([x * (1.0 + 0.02 * i) for i in range(n_steps)])                  # every rank: its own data shard
    ms = (xs[:, 0] > 300.0).long()
    mk = lambda lo, hi: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xs[lo:hi], ms[lo:hi]), batch_size=a.batch)

    model = S.SwinWNet(error_matrix=True, depths=[2, 2, 2, 2])
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    sync = S.train.DistributedGradSync(model) if world > 1 else None
    tr = FullModelTrainer(model, mk(0, 1), mk(0, 1), dev, num_epochs=100, warmup_epochs=1, lr=1e-4, verbose=False)
    tr.optimizer = S.train.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    tr.scheduler = tr._build_default_scheduler()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # first pair: losses, next to the reference model on the same batches (rank 0)
    tr.train_loader = mk(0, 2 * a.batch)
    first = tr._run_epoch(0, train=True)
    reduced = sync.reduced_elements if sync else 0
    ref_first = None
    if rank == 0:
        ref = R.SwinWNet(error_matrix=True, depths=[2, 2, 2, 2])
        ref.load_state_dict(sd, strict=True)
        rt = FullModelTrainer(ref.to(dev), mk(0, 2 * a.batch), mk(0, 1), dev, num_epochs=100, warmup_epochs=1, lr=1e-4, verbose=False)
        ref_first = rt._run_epoch(0, train=True)
        del rt, ref
        torch.cuda.empty_cache()
    for w in range(1, a.warmup):
        tr.train_loader = mk(2 * w * a.batch, 2 * (w + 1) * a.batch)
        tr._run_epoch(0, train=True)
    tr.train_loader = mk(2 * a.warmup * a.batch, n_steps * a.batch)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = tr._run_epoch(0, train=True)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        steps = 2 * a.pairs
        line = {"metric": "FullModelTrainer even+odd step pairs, training diffractions/sec", "unit": "diffractions/s",
                "value": world * a.batch * steps / (t.item() / 1e3), "n_gpus": world, "steps": steps, "warmup": 2 * a.warmup,
                "ms_per_step": t.item() / steps, "higher_is_better": True, "scaling": "weak", "dtype": "fp16 forward kernels, fp32 ATen backward",
                "data": "synthetic", "config": {"workload": "configs[4]: FullModel_supervised_trainer even+odd steps, [B,1,250,480], depths [2,2,2,2]",
                                                "batch_per_gpu": a.batch, "global_batch": a.batch * world, "parallelism": f"dp{world}"},
                "first_pair_losses": first, "reference_first_pair_losses": ref_first, "last_epoch_losses": last,
                "allreduced_elements_per_backward": reduced, "replicas_identical": bool(lo.item() == hi.item()),
                "backward": "KernelOp recompute through ATen (no hand-written backward kernels yet)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
