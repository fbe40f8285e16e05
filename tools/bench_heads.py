"""Micro-benchmark / ncu target of the conv heads and glue kernels at the model's shapes (batch 64):
recon head (500x960, 12 ch), segmentation head (125x240 tokens -> 250x480 and 500x960), patch embed, copy_cols."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swinwnet_b200 as S  # noqa: E402
from swinwnet_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--once", action="store_true")
ap.add_argument("--ops", default="recon,seg,embed")
a = ap.parse_args()
DEV, B = "cuda", a.batch


def timeit(name, fn, byts):
    fn()
    torch.cuda.synchronize()
    if a.once:
        fn()
        torch.cuda.synchronize()
        return
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"{name:12s} {ms:8.3f} ms  {byts / ms / 1e6:8.1f} GB/s", flush=True)


if "recon" in a.ops:
    tok = torch.randn(B, 500 * 960, 12, device=DEV)
    w1, b1, w2, b2 = (torch.randn(12, 12, 3, 3, device=DEV) * 0.1, torch.zeros(12, device=DEV), torch.randn(2, 12, 1, 1, device=DEV) * 0.3,
                      torch.zeros(2, device=DEV))
    out = torch.empty(B, 2, 500, 960, device=DEV)
    timeit("recon_head", lambda: ops.recon_head(tok, w1, b1, w2, b2, out, B, 500, 960, 2, 500, 960), tok.numel() * 4 + out.numel() * 4)
    del tok, out
if "seg" in a.ops:
    tok = torch.randn(B, 125 * 240, 48, device=DEV)
    w1, b1, w2, b2 = (torch.randn(24, 48, 3, 3, device=DEV) * 0.05, torch.zeros(24, device=DEV), torch.randn(1, 24, 1, 1, device=DEV) * 0.2,
                      torch.zeros(1, device=DEV))
    for up in (2, 4):
        low = torch.empty(B, 125, 240, device=DEV)
        out = torch.empty(B, 1, 125 * up, 240 * up, device=DEV)
        timeit(f"seg_head x{up}", lambda: ops.seg_head(tok, w1, b1, w2, b2, low, out, B, 125, 240, up, 125 * up, 240 * up),
               tok.numel() * 4 + out.numel() * 4)
if "embed" in a.ops:
    for (H, W, s) in ((250, 480, 1), (500, 960, 2)):
        x = torch.randn(B, 2, H, W, device=DEV)
        w, b = torch.randn(48, 2, 2, 2, device=DEV), torch.zeros(48, device=DEV)
        out = torch.empty(B, 125 * 240, 48, device=DEV)
        timeit(f"embed s{s}", lambda: ops.patch_embed(x, w, b, torch.ones(48, device=DEV), torch.zeros(48, device=DEV), out, B, 2, H, W, 125, 240, s),
               x.numel() * 4 // (s * s) + out.numel() * 4)
