"""Per-op micro-benchmark at the model's (tokens, channels) shapes for batch B: CUDA-event time, achieved
algorithmic GB/s and TFLOP/s per kernel.  Usage: python tools/bench_ops.py [--batch 64] [--ops mlp,qkv,proj,attn]
[--only C] [--reps 5].  With --once each op runs exactly once after one warm-up (ncu target)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swinwnet_b200 as S  # noqa: E402
from swinwnet_b200 import ops, packing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--ops", default="mlp,qkv,proj,attn")
ap.add_argument("--only", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--once", action="store_true")
ap.add_argument("--inplace", action="store_true", help="out aliases x (as the model runs blocks)")
a = ap.parse_args()
DEV = "cuda"
torch.manual_seed(0)
OPD = S.ops.operand_dtype()   # 16-bit tensor-core operand dtype of the built library (fp16 default)
# (tokens per image, C, heads, grid H, W)
SHAPES = [(30000, 48, 3, 125, 240), (7560, 96, 6, 63, 120), (1920, 192, 12, 32, 60), (480, 384, 24, 16, 30),
          (1920, 384, 12, 32, 60), (7560, 192, 6, 63, 120), (30000, 96, 3, 125, 240), (120000, 24, 3, 250, 480),
          (480000, 12, 3, 500, 960)]


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    if a.once:
        fn()
        torch.cuda.synchronize()
        return 0.0
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def report(name, M, C, ms, byts, flops):
    if ms > 0:
        print(f"{name:6s} M={M:9d} C={C:4d}  {ms:8.3f} ms  {byts / ms / 1e6:8.1f} GB/s  {flops / ms / 1e9:8.2f} TFLOP/s", flush=True)


for (L, C, nH, H, W) in SHAPES:
    if a.only and C != a.only:
        continue
    B = a.batch
    M = B * L
    x = torch.randn(M, C, device=DEV)
    lw, lb = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    if "blk" in a.ops and C in (12, 24):
        shp = [(C,), (C,), (3 * C, C), (3 * C,), (81, nH), (C, C), (C,), (C,), (C,), (4 * C, C), (4 * C,), (C, 4 * C), (C,)]
        params = [torch.randn(*s, device=DEV) * 0.1 for s in shp]
        out = torch.empty_like(x)
        ms = timeit(lambda: ops.swin_block_small(x, out, B, H, W, C, nH, 0, 1e-5, params))
        report("blk", M, C, ms, M * C * 8, (24.0 * C * C + 100.0 * C) * M)
    if "fused" in a.ops and C <= 48:
        shp = [(C,), (C,), (3 * C, C), (3 * C,), (81, nH), (C, C), (C,), (C,), (C,), (4 * C, C), (4 * C,), (C, 4 * C), (C,)]
        params = [torch.randn(*s, device=DEV) * 0.1 for s in shp]
        Wpk, fpk = packing.pack_fused_block(*params, nH)
        out = torch.empty_like(x)
        for do_mlp in (True, False):
            ms = timeit(lambda: ops.swin_block_fused(x, out, B, H, W, C, nH, 1e-5, Wpk, fpk, do_mlp))
            report("fblk" if do_mlp else "fattn", M, C, ms, M * C * 8, ((24.0 if do_mlp else 8.0) * C * C + 100.0 * C) * M)
    if "warp" in a.ops and C in (12, 24, 48):
        shp = [(C,), (C,), (3 * C, C), (3 * C,), (81, nH), (C, C), (C,), (C,), (C,), (4 * C, C), (4 * C,), (C, 4 * C), (C,)]
        params = [torch.randn(*s, device=DEV) * 0.1 for s in shp]
        Wpk, fpk = packing.pack_warp_block(*params, nH)
        out = torch.empty_like(x)
        ms = timeit(lambda: ops.swin_block_warp(x, out, B, H, W, C, nH, 1e-5, Wpk, fpk))
        report("wblk", M, C, ms, M * C * 8, (24.0 * C * C + 100.0 * C) * M)
        W2, f2 = torch.cat([Wpk, Wpk]), torch.cat([fpk, fpk])
        ms = timeit(lambda: ops.swin_block_warp(x, out, B, H, W, C, nH, 1e-5, W2, f2, 2))
        report("wblk2", M, C, ms, M * C * 8, 2 * (24.0 * C * C + 100.0 * C) * M)
    if "fused" in a.ops and C == 96:
        shp = [(C,), (C,), (3 * C, C), (3 * C,), (81, nH), (C, C), (C,)]
        params = [torch.randn(*s, device=DEV) * 0.1 for s in shp]
        Wpk, fpk = packing.pack_fused_attn_stream(*params, nH)
        out = torch.empty_like(x)
        ms = timeit(lambda: ops.swin_block_fused(x, out, B, H, W, C, nH, 1e-5, Wpk, fpk, False))
        report("fattn", M, C, ms, M * C * 8, (8.0 * C * C + 100.0 * C) * M)
    if "mlp" in a.ops:
        W1, b1 = torch.randn(4 * C, C, device=DEV) * C ** -0.5, torch.zeros(4 * C, device=DEV)
        W2, b2 = torch.randn(C, 4 * C, device=DEV) * (4 * C) ** -0.5, torch.zeros(C, device=DEV)
        HC, TR = ops.mlp_config(C)
        Wp, b2p = packing.pack_mlp(W1, W2, b2, HC, TR)
        out = x if a.inplace else torch.empty_like(x)
        ms = timeit(lambda: ops.mlp(x, out, M, C, lw, lb, Wp, b1, b2p))
        report("mlp", M, C, ms, M * C * 8, 16.0 * M * C * C)
    if "qkv" in a.ops:
        Wq, bq = torch.randn(3 * C, C, device=DEV) * C ** -0.5, torch.zeros(3 * C, device=DEV)
        nv = packing.choose_chunk(3 * C, int(os.environ.get("SWN_QKV_NT", 256)))
        Wp, bp, NT, nch = packing.pack_rowgemm(Wq, bq, nv)
        qkv = torch.empty(M, 3 * C, device=DEV, dtype=OPD)
        ms = timeit(lambda: ops.rowgemm(A=x, a_mode=ops.A_F32_LN, M=M, K=C, lda=C, ln_w=lw, ln_b=lb, Wp=Wp, NT=NT, nchunks=nch,
                                        n_valid=nv, e_mode=ops.E_BF16, bias=bp, out=qkv, ldo=3 * C))
        report("qkv", M, C, ms, M * C * 10, 6.0 * M * C * C)
    if "attn" in a.ops:
        qkv = torch.randn(M, 3 * C, device=DEV).to(OPD)
        att = torch.empty(M, C, device=DEV, dtype=OPD)
        bq, tab = torch.zeros(3 * C, device=DEV), torch.zeros(81, nH, device=DEV)
        ms = timeit(lambda: ops.window_attention(qkv, att, bq, tab, B, H, W, C, nH, 0))
        report("attn", M, C, ms, M * C * 8, 100.0 * M * C)
        fr = packing.rel_pos_bias_fragments(tab, 1.4426950408889634).reshape(-1).contiguous()
        ms = timeit(lambda: ops.window_attention(qkv, att, bq, tab, B, H, W, C, nH, 0, fr))
        report("attnf", M, C, ms, M * C * 8, 100.0 * M * C)
    if "proj" in a.ops:
        att = torch.randn(M, C, device=DEV).to(OPD)
        Wo, bo = torch.randn(C, C, device=DEV) * C ** -0.5, torch.zeros(C, device=DEV)
        nv = packing.choose_chunk(C, int(os.environ.get("SWN_PROJ_NT", 256)))
        Wp, bp, NT, nch = packing.pack_rowgemm(Wo, bo, nv)
        out = x if a.inplace else torch.empty_like(x)
        ms = timeit(lambda: ops.rowgemm(A=att, a_mode=ops.A_BF16, M=M, K=C, lda=C, Wp=Wp, NT=NT, nchunks=nch, n_valid=nv,
                                        e_mode=ops.E_F32, bias=bp, out=out, ldo=C, res=x, ldres=C))
        report("proj", M, C, ms, M * C * 10, 2.0 * M * C * C)
    if "merge" in a.ops and C <= 192:
        pm = S.model.PatchMerging(C).to(DEV)
        xm = x.view(B, L, C)
        with torch.no_grad():
            ms = timeit(lambda: pm(xm, (H, W)))
        Mo = B * ((H + 1) // 2) * ((W + 1) // 2)
        report("merge", M, C, ms, M * C * 4 + Mo * 2 * C * 4, 2.0 * Mo * 4 * C * 2 * C)
    if "expand" in a.ops and C >= 24:
        pe = S.model.PatchExpanding(C).to(DEV)
        xe = x.view(B, L, C)
        with torch.no_grad():
            ms = timeit(lambda: pe.run(xe, (H, W)))
        report("expand", M, C, ms, M * C * 4 + M * 4 * (C // 2) * 4, 2.0 * M * C * 2 * C)
    if "declin" in a.ops and C >= 96:
        Wl, bl = torch.randn(C // 2, C, device=DEV) * C ** -0.5, torch.zeros(C // 2, device=DEV)
        Wp, bp, NT, nch = packing.pack_rowgemm(Wl, bl, packing.choose_chunk(C // 2, 256))
        o = torch.empty(M, C // 2, device=DEV)
        ms = timeit(lambda: ops.rowgemm(A=x, a_mode=ops.A_F32, M=M, K=C, lda=C, Wp=Wp, NT=NT, nchunks=nch, n_valid=packing.choose_chunk(C // 2, 256),
                                        e_mode=ops.E_F32, bias=bp, out=o, ldo=C // 2))
        report("declin", M, C, ms, M * C * 6, 2.0 * M * C * (C // 2))
    del x
    torch.cuda.empty_cache()
