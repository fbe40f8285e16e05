"""Where the cycles of mlp_kernel go (clock64 sums over all CTAs): waits of the MMA-issuing warp and of one epilogue warp.
Needs a profiling build: SWN_NVCC_EXTRA=-DSWN_MLP_PROFILE=1 python __graft_entry__.py build (the product build compiles the
clocks out).  Usage: python tools/mlp_phase_profile.py [--C 192] [--M 483840]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swinwnet_b200 as S  # noqa: E402
from swinwnet_b200 import ops, packing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=192)
ap.add_argument("--M", type=int, default=483840)
a = ap.parse_args()
C, M = a.C, a.M
x = torch.randn(M, C, device="cuda")
out = torch.empty_like(x)
lw, lb = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
W1, b1 = torch.randn(4 * C, C, device="cuda") * C ** -0.5, torch.zeros(4 * C, device="cuda")
W2, b2 = torch.randn(C, 4 * C, device="cuda") * (4 * C) ** -0.5, torch.zeros(C, device="cuda")
HC, TR = ops.mlp_config(C)
Wp, b2p = packing.pack_mlp(W1, W2, b2, HC, TR)
for _ in range(2):
    ops.mlp(x, out, M, C, lw, lb, Wp, b1, b2p)
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
ops.set_phase_profile(buf)
ops.mlp(x, out, M, C, lw, lb, Wp, b1, b2p)
torch.cuda.synchronize()
ops.set_phase_profile(None)
t = buf.cpu().tolist()
ntiles = (M + 127) // 128
names = ["MMA: wait A tile", "MMA: wait W1 tiles", "MMA: wait Hacc free", "MMA: wait hidden tile", "MMA: wait W2 tiles", "MMA warp total",
         "EPI: wait Hacc full", "EPI: wait hidden free", "EPI: GELU chunks", "EPI: wait Y", "EPI: final epilogue", "CTA total",
         "prologue (LN -> A) | persist: LN wait rows", "persist: LN wait A free", "persist: LN compute", ""]
if C <= 96:
    names[5] = "MMA: wait Y drained"
print(f"C={C} M={M} HC={HC} TR={TR} tiles={ntiles}")
for i, n in enumerate(names):
    print(f"{n:24s} {t[i] / ntiles:10.0f} cycles / tile")
