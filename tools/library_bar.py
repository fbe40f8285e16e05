"""The library bar: the UNMODIFIED reference (baseline/_ref) on the same B200 through its
own public API — `SwinWNetInference(model, device)(images)` — in eager fp32, TF32 and bf16 / fp16 autocast, i.e. what
cuBLAS + ATen deliver for this path without any of this repo's kernels.  CUDA events, 3 warm-up calls, median of N.

    python tools/library_bar.py [--batches 1 8 64] [--out gpurun_out/r2_ref_eager_b200.json]
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


def measure(batches=(1, 8, 64), modes=("fp32", "tf32", "bf16_autocast", "fp16_autocast"), iters=5, device="cuda:0", verbose=True):
    ref_dir, R, P = import_reference()
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    model = R.SwinWNet(error_matrix=True, depths=[2, 2, 2, 2])
    model.load_state_dict(benchdata.make_state_dict(man["wnet_em"], seed=1), strict=True)
    inf = P.SwinWNetInference(model, device)
    res = {"reference_dir": ref_dir, "device": torch.cuda.get_device_name(0), "torch": torch.__version__, "runs": []}
    for B in batches:
        base = benchdata.synthetic_diffractions(min(B, 8), seed=100, two_channel=False)
        x = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:B].contiguous().to(device)
        for mode in modes:
            tf32 = mode != "fp32"
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            ctx = (torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16_autocast" else
                   torch.autocast("cuda", dtype=torch.float16) if mode == "fp16_autocast" else torch.autocast("cuda", enabled=False))
            try:
                with ctx:
                    for _ in range(3):
                        inf(x)
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(iters):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        inf(x)
                        e1.record()
                        torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                ts.sort()
                ms = ts[len(ts) // 2]
                res["runs"].append({"batch": B, "mode": mode, "ms_per_call": ms, "diffractions_per_s": B / ms * 1e3,
                                    "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
            except Exception as e:  # e.g. out of memory at the largest batch in fp32
                res["runs"].append({"batch": B, "mode": mode, "error": f"{type(e).__name__}: {str(e)[:200]}"})
            torch.cuda.reset_peak_memory_stats()
            torch.cuda.empty_cache()
            if verbose:
                print(res["runs"][-1], flush=True)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = True
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 8, 64])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_ref_eager_b200.json"))
    a = ap.parse_args()
    r = measure(tuple(a.batches), iters=a.iters)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(r, open(a.out, "w"), indent=1)
