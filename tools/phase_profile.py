"""Per-phase cycle breakdown of the fused block kernels (thread 0 of every CTA, clock64 between CTA barriers).
Usage: python tools/phase_profile.py [--C 96] [--heads 3] [--batch 64]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swinwnet_b200 as S  # noqa: E402
from swinwnet_b200 import ops, packing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=96)
ap.add_argument("--heads", type=int, default=3)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--H", type=int, default=125)
ap.add_argument("--W", type=int, default=240)
ap.add_argument("--attn-only", action="store_true")
a = ap.parse_args()
C, nH, B, H, W = a.C, a.heads, a.batch, a.H, a.W
M = B * H * W
x = torch.randn(M, C, device="cuda")
out = torch.empty_like(x)
shp = [(C,), (C,), (3 * C, C), (3 * C,), (81, nH), (C, C), (C,), (C,), (C,), (4 * C, C), (4 * C,), (C, 4 * C), (C,)]
params = [torch.randn(*s, device="cuda") * 0.1 for s in shp]
if C == 96:
    Wpk, fpk = packing.pack_fused_attn_stream(*params[:7], nH)
    do_mlp = False
    names = ["load wait", "LN stats", "LN normalise", "qkv MMA", "qkv epilogue", "attention", "proj: barrier", "epilogue",
             "proj: MMA issue", "proj: issue loads", "proj: MMA wait", "(unused)",
             "qkv: weight tile 0 landed", "qkv: tile 1 landed", "qkv: tile 2 landed", "qkv: tile 3 landed"]
else:
    Wpk, fpk = packing.pack_fused_block(*params, nH)
    do_mlp = not a.attn_only
    names = ["write-back(prev)+tok", "loads landed", "LN1 stats", "LN1 normalise", "qkv MMA", "qkv epilogue", "attention",
             "proj MMA", "proj epilogue", "LN2", "fc1 MMA", "GELU (batched)", "hidden-tile wait", "GELU (chunk)", "fc2 MMA",
             "final epilogue"]
for _ in range(2):
    ops.swin_block_fused(x, out, B, H, W, C, nH, 1e-5, Wpk, fpk, do_mlp)
buf = torch.zeros(148 * 4, 16, dtype=torch.int64, device="cuda")
ops.set_phase_profile(buf)
ops.swin_block_fused(x, out, B, H, W, C, nH, 1e-5, Wpk, fpk, do_mlp)
torch.cuda.synchronize()
ops.set_phase_profile(None)
t = buf.double().sum(0).cpu()
tot = t.sum().item()
for i, n in enumerate(names):
    if t[i] > 0:
        print(f"{n:16s} {100 * t[i].item() / tot:5.1f}%   {t[i].item() / max((buf[:, i] > 0).sum().item(), 1):12.0f} cycles / CTA")
print(f"total {tot / max((buf.sum(1) > 0).sum().item(), 1):.0f} cycles / CTA")
