#!/usr/bin/env python
"""Benchmark of the SwinWNet forward hot path (BASELINE.json metric: diffractions/s).

A "step" is one pass of the full ST inference pipeline (segment_1 -> upscale -> segment_2 + glue,
ST_Inference_Pipline.py:73-136) over one batch of synthetic diffractions of the dataset shape
[B,1,250,480] (error channel derived -> multimodal [B,2,250,480] model, BASELINE configs[1], batch 64
per GPU).  N>1: every rank runs its own batch shard (weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = pinned host inputs + D2H of the
result inside the timed region, through the public SwinWNetInference call.  `--impl reference` times the
oracle (CPU port of the reference, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 250, 480
DEPTHS = [2, 2, 2, 2]
GFLOP_PER_DIFFRACTION = 207.76          # reference-convention algorithmic FLOPs
METRIC, UNIT = "SwinWNet fwd diffractions/sec", "diffractions/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def mlp_flops_per_diffraction():
    """16*M*C^2 per fused-MLP launch (fc1 + fc2), summed over the 52 blocks of one pipeline pass."""
    enc = [(30000, 48), (7560, 96), (1920, 192), (480, 384)]
    dec = [(1920, 384), (7560, 192), (30000, 96)]
    per_pass = sum(2 * 16 * m * c * c for m, c in enc + [(480, 384)] + dec)        # depth 2 each
    sr_head = sum(2 * 16 * m * c * c for m, c in [(120000, 24), (480000, 12)])
    return 3 * per_pass + sr_head


def shard_bounds(n_items, rank, world):
    """contiguous shard [lo, hi) of `n_items` independent diffractions owned by `rank` (inference shards by batch;
    no data-path collective)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, dist=None, device="cpu"):
    """the job's time is the slowest rank's time: all-reduce MAX (NCCL on the GPU box, gloo in the CPU tests)."""
    t = torch.tensor([float(value)], device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def whole_job_rate(units_per_rank, world, steps, ms):
    """aggregate throughput over all ranks (weak scaling: every rank processes `units_per_rank` per step)."""
    return world * units_per_rank * steps / (ms / 1e3)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def run_reference(args, rank):
    """the reference's own algorithm on the host CPU (oracle port; the Python reference cannot travel)."""
    if rank != 0:
        return
    from oracle import swinwnet_oracle as O
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    sd = O.make_state_dict(man["wnet_em"], seed=1)
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    sample_b = 1                                                  # bounded sample: 1 diffraction per step
    x = O.synthetic_diffractions(sample_b, seed=0, two_channel=False)
    steps, warm = min(args.steps, 6), min(args.warmup, 1)
    with torch.no_grad():
        for _ in range(warm):
            O.st_pipeline(sd, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.st_pipeline(sd, x)
        dt = (time.perf_counter() - t0) / steps
    v = sample_b / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SwinWNet multimodal ST pipeline [B,2,250,480], depths [2,2,2,2], random-init weights",
                       "sample_batch": sample_b},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{steps} steps x {sample_b} diffraction, fp32 torch CPU oracle"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="diffractions per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import swinwnet_b200 as S
    from swinwnet_b200 import ops
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    import benchdata                               # seeded synthetic inputs / weights (no model math)
    model = S.SwinWNet(error_matrix=True, depths=DEPTHS)
    model.load_state_dict(benchdata.make_state_dict(man["wnet_em"], seed=1), strict=True)
    inf = S.SwinWNetInference(model, dev, max_batch=64)
    B = args.batch
    base = benchdata.synthetic_diffractions(min(B, 8), seed=100 + rank, two_channel=False)
    x_host = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:B].contiguous()
    x_host = (x_host * (1.0 + 0.01 * torch.arange(B).view(B, 1, 1, 1))).pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty(B, 2, 2 * H, 2 * W).pin_memory()
    W_ = max(args.warmup, 3)

    # ---- fused-MLP kernel timing hook (the dominant kernel: 57 % of the FLOPs) ----
    mlp_events = []
    orig_mlp = ops.mlp

    def timed_mlp(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_mlp(*a, **k)
        e1.record()
        mlp_events.append((e0, e1))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dist, dev)

    def step_dev():
        inf(x_dev)

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        out = inf(xd)
        out_host.copy_(out, non_blocking=True)

    for _ in range(W_):
        step_dev()
    sampler = ClockSampler(local)
    sampler.start()
    ops.mlp = timed_mlp
    S.model.ops.mlp = timed_mlp
    n0 = ops.LAUNCH_COUNT
    ms = timed(step_dev, args.steps)
    launches = ops.LAUNCH_COUNT - n0
    ops.mlp = orig_mlp
    S.model.ops.mlp = orig_mlp
    mlp_ms = sum(a.elapsed_time(b) for a, b in mlp_events) / args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=3)

    value = whole_job_rate(B, world, args.steps, ms)
    e2e_v = whole_job_rate(B, world, args.steps, ms_e2e)
    pk = peaks()
    mlp_tflops = mlp_flops_per_diffraction() * B / (mlp_ms / 1e3) / 1e12
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W_,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if ops.operand_dtype() == torch.bfloat16 else "fp16", "data": "synthetic",
            "config": {"workload": "SwinWNet multimodal ST pipeline [B,2,250,480] (configs[1]), depths [2,2,2,2], "
                                   "random-init weights", "batch_per_gpu": B, "global_batch": B * world,
                       "l2_policy": "per-step working set (>2 GB of activations) exceeds the 126 MB L2"},
            "model_tflops": GFLOP_PER_DIFFRACTION * value / 1e3,
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "swn::mlp_kernel (fused LN+fc1+GELU+fc2+residual)",
                         "achieved": mlp_tflops, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": mlp_tflops / pk["tf_sust"],
                         "traffic": None, "peak_source": pk["src"] + " (sustained bf16)",
                         "kernel_ms_per_step": mlp_ms, "kernel_share_of_step": mlp_ms / (ms / args.steps)},
            "clocks": sampler.summary()}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        from oracle import swinwnet_oracle as O    # the checker, executed only for the reported CPU baseline
        sd = benchdata.make_state_dict(man["wnet_em"], seed=1)
        xs = x_host[:1].clone()
        with torch.no_grad():
            O.st_pipeline(sd, xs)
            t0 = time.perf_counter()
            n = 4
            for _ in range(n):
                O.st_pipeline(sd, xs)
            dt = (time.perf_counter() - t0) / n
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n} x 1 diffraction of the same workload, fp32 torch CPU oracle"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
