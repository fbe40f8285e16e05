"""Profiling target: `--iters` full ST-pipeline calls at the bench workload (no timing, no oracle compute)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swinwnet_b200 as S  # noqa: E402
import benchdata as O  # noqa: E402  (seeded weight / input generator, no model math)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
m = S.SwinWNet(error_matrix=True, depths=[2, 2, 2, 2])
m.load_state_dict(O.make_state_dict(man["wnet_em"], seed=1), strict=True)
inf = S.SwinWNetInference(m, "cuda:0")
x = O.synthetic_diffractions(min(a.batch, 4), seed=1, two_channel=False)
x = x.repeat((a.batch + 3) // 4, 1, 1, 1)[:a.batch].to("cuda:0")
for _ in range(a.iters):
    inf(x)
torch.cuda.synchronize()
print("ok", S.ops.LAUNCH_COUNT)
