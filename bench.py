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


# dram__bytes_read.sum + dram__bytes_write.sum per step of the kernel family, summed over its launches, from the
# `ncu --set full` captures summarised under profiles/ (null = not captured for this build)
TRAFFIC_NCU = {"fused": 22.7e9, "mlp": None}   # bytes per step (22 launches at batch 64), profiles/r1_ncu_fused_*.txt


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
    ap.add_argument("--graph", action="store_true", help="also time CUDA-graph replay of the pipeline (matters at small --batch, where a pass is launch-bound); reported under \"graph\"")
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
    inf_fast = S.SwinWNetInference(model, dev, max_batch=64, cuda_graph=True) if args.graph else inf
    B = args.batch
    base = benchdata.synthetic_diffractions(min(B, 8), seed=100 + rank, two_channel=False)
    x_host = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:B].contiguous()
    x_host = (x_host * (1.0 + 0.01 * torch.arange(B).view(B, 1, 1, 1))).pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty(B, 2, 2 * H, 2 * W).pin_memory()
    W_ = max(args.warmup, 3)

    # ---- per-kernel-family timing hooks (CUDA events on the launching stream, inside the timed region) ----
    # fused = swn::swin_fused_kernel / swin_attn_stream_kernel (W-MSA [+ MLP] in one tcgen05 kernel, C <= 96),
    # mlp   = swn::mlp_kernel / mlp_persist_kernel (LN + fc1 + GELU + fc2 + residual, C >= 96)
    fam = {"fused": {"ev": [], "flop": 0.0, "bytes": 0.0}, "mlp": {"ev": [], "flop": 0.0, "bytes": 0.0}}
    orig = {"fused": ops.swin_block_fused, "mlp": ops.mlp}

    def hook(name, work):
        fn = orig[name]

        def timed_op(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*a, **k)
            e1.record()
            fl, by = work(*a, **k)
            fam[name]["ev"].append((e0, e1))
            fam[name]["flop"] += fl
            fam[name]["bytes"] += by
        return timed_op

    def fused_work(x, out, Bn, Hn, Wn, C, nH, eps, Wpk, fpk, do_mlp=True):
        M = Bn * Hn * Wn      # algorithmic FLOPs of a block: 24 C^2 + 100 C per token (8 C^2 + 100 C for the W-MSA half)
        return ((24.0 if do_mlp else 8.0) * C * C + 100.0 * C) * M, 8.0 * M * C

    def mlp_work(x, out, M, C, *a, **k):
        return 16.0 * M * C * C, 8.0 * M * C
    hooks = {"fused": hook("fused", fused_work), "mlp": hook("mlp", mlp_work)}

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

    # one chunk per step: the H2D of step i+1 and the D2H of step i then overlap the compute of their neighbours across
    # steps (copy streams), all inside the timed region; smaller chunks also overlap inside a step but run the deep,
    # narrow layers of the net on half-filled waves
    E2E_CHUNK = int(os.environ.get("SWN_E2E_CHUNK", B))

    def step_e2e():
        # public host-data call: pinned host inputs -> pinned host result, copies pipelined against compute in chunks
        inf.run_host(x_host, out=out_host, chunk=E2E_CHUNK)

    for _ in range(W_):
        step_dev()
    sampler = ClockSampler(local)
    sampler.start()
    ops.swin_block_fused, ops.mlp = hooks["fused"], hooks["mlp"]
    n0 = ops.LAUNCH_COUNT
    ms = timed(step_dev, args.steps)
    launches = ops.LAUNCH_COUNT - n0
    ops.swin_block_fused, ops.mlp = orig["fused"], orig["mlp"]
    pk = peaks()
    kern = {}
    for name, label in (("fused", "swn::swin_fused_kernel + swin_attn_stream_kernel (fused LN1+qkv+W-MSA+proj[+LN2+MLP], C<=96)"),
                        ("mlp", "swn::mlp_kernel + mlp_persist_kernel (fused LN2+fc1+GELU+fc2+residual, C>=96)")):
        f = fam[name]
        k_ms = sum(a.elapsed_time(b) for a, b in f["ev"]) / args.steps
        tf = f["flop"] / args.steps / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
        gbs = f["bytes"] / args.steps / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
        kern[name] = {"kernel": label, "launches_per_step": len(f["ev"]) // max(args.steps, 1), "kernel_ms_per_step": k_ms,
                      "kernel_share_of_step": k_ms / (ms / args.steps), "achieved_tflops": tf, "tensor_frac": tf / pk["tf_sust"],
                      "achieved_gbs": gbs, "hbm_frac": gbs / pk["hbm"]}
    dom = max(kern, key=lambda n: kern[n]["kernel_ms_per_step"])
    # which roof bounds the family: its algorithmic intensity (FLOP per algorithmic byte) against the ridge of the two
    # measured peaks.  The fused W-MSA / block family sits below the ridge (3C + 12.5 FLOP/B per token-block, C <= 96).
    kd = kern[dom]
    ridge = pk["tf_sust"] * 1e12 / (pk["hbm"] * 1e9)
    intensity = fam[dom]["flop"] / max(fam[dom]["bytes"], 1.0)
    hbm_bound = intensity < ridge
    roof = {"bound": "hbm" if hbm_bound else "tensor", "kernel": kd["kernel"],
            "achieved": kd["achieved_gbs"] if hbm_bound else kd["achieved_tflops"],
            "peak": pk["hbm"] if hbm_bound else pk["tf_sust"], "unit": "GB/s" if hbm_bound else "TFLOP/s",
            "frac": kd["hbm_frac"] if hbm_bound else kd["tensor_frac"], "traffic": TRAFFIC_NCU.get(dom),
            "peak_source": pk["src"] + (" (copy bandwidth)" if hbm_bound else " (sustained bf16)"),
            "intensity_flop_per_byte": intensity, "ridge_flop_per_byte": ridge,
            "kernel_ms_per_step": kd["kernel_ms_per_step"], "kernel_share_of_step": kd["kernel_share_of_step"],
            "tensor_tflops": kd["achieved_tflops"], "tensor_frac": kd["tensor_frac"],
            "hbm_gbs": kd["achieved_gbs"], "hbm_frac": kd["hbm_frac"]}
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=3)

    graph_info = None
    if args.graph:
        for _ in range(3):
            inf_fast(x_dev)
        ms_g = timed(lambda: inf_fast(x_dev), args.steps)
        graph_info = {"value": whole_job_rate(B, world, args.steps, ms_g), "unit": UNIT, "ms_per_step": ms_g / args.steps}
    value = whole_job_rate(B, world, args.steps, ms)
    e2e_v = whole_job_rate(B, world, args.steps, ms_e2e)
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
            "roofline": roof,
            "kernels": kern,
            "clocks": sampler.summary()}
    if graph_info:
        line["graph"] = graph_info
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
