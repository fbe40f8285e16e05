"""BASELINE.json configs[2]: the single-branch models at batch 256 (parity-test cases in bench.py's contract, timed here
for the record).  SwinUNet [256,1,250,480] -> [256,1,250,480] (16.17 TFLOP per call) and SwinUNetSR
[256,1,250,480] -> [256,1,500,960] (18.62 TFLOP), micro-batched 64 at a time, synthetic normalised inputs, seeded
random-init weights.  Usage: python tools/bench_config3.py [--batch 256] [--reps 5]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swinwnet_b200 as S  # noqa: E402
import benchdata  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--micro", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
x = benchdata.synthetic_diffractions(8, seed=3, two_channel=False)
x = (x / x.amax(dim=(2, 3), keepdim=True)).repeat(a.batch // 8 + 1, 1, 1, 1)[:a.batch].cuda()
for name, cls, key, gflop in (("SwinUNet", S.SwinUNet, "unet", 63.18), ("SwinUNetSR", S.SwinUNetSR, "unetsr", 72.73)):
    m = cls(depths=[2, 2, 2, 2])
    m.load_state_dict(benchdata.make_state_dict(man[key], seed=1), strict=True)
    m = m.cuda().eval()

    def run():
        with torch.no_grad():
            return [m(x[i:i + a.micro]) for i in range(0, a.batch, a.micro)]
    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:10s} batch {a.batch}: {ms:8.2f} ms  {a.batch / ms * 1e3:8.1f} diffractions/s  {gflop * a.batch / ms:7.1f} TFLOP/s (algorithmic)  "
          f"out {tuple(out[0].shape[1:])}")
