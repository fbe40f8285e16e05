"""Diagnostic (not a test): runs small tcgen05 rowgemm / mlp cases and prints structured error information so a
descriptor / layout bug can be localised from one GPU run."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swinwnet_b200 as S  # noqa: E402
from swinwnet_b200 import ops, packing  # noqa: E402

torch.manual_seed(0)
DEV = "cuda"
OPD = S.ops.operand_dtype()   # 16-bit tensor-core operand dtype of the built library (fp16 default)


def report(name, out, ref):
    out, ref = out.float().cpu(), ref.float().cpu()
    err = (out - ref).abs()
    print(f"[{name}] max|ref|={ref.abs().max():.3f} max err={err.max():.4f} finite={bool(torch.isfinite(out).all())}")
    if err.max() > 0.05 * ref.abs().max():
        M, N = ref.shape
        rb = err.view(M // 8, 8, N).amax(dim=(1, 2))
        print("   err by 8-row block :", [round(v, 2) for v in rb.tolist()][:16])
        cb = err.view(M, -1, 8).amax(dim=(0, 2)) if N % 8 == 0 else err.amax(0)
        print("   err by 8-col block :", [round(v, 2) for v in cb.tolist()][:32])
        # does each output column match SOME reference column (permutation)?
        a = out / (out.norm(dim=0, keepdim=True) + 1e-6)
        b = ref / (ref.norm(dim=0, keepdim=True) + 1e-6)
        corr = (a.t() @ b)
        best = corr.abs().max(dim=1)
        print("   best-match ref col per out col:", best.indices.tolist()[:32])
        print("   match quality              :", [round(v, 2) for v in best.values.tolist()][:32])
        print("   out[0,:8]", out[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())


def gemm_case(M, K, N, nv=None):
    nv = nv or packing.choose_chunk(N, 256)
    a = torch.randn(M, K).to(OPD)
    W = torch.randn(N, K) * K ** -0.5
    ref = a.float() @ W.to(OPD).float().t()
    Wp, _, NT, nch = packing.pack_rowgemm(W.to(DEV), None, nv)
    out = torch.zeros(M, N, device=DEV)
    ops.rowgemm(A=a.to(DEV), a_mode=ops.A_BF16, M=M, K=K, lda=K, Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_F32,
                out=out, ldo=N)
    torch.cuda.synchronize()
    report(f"gemm M{M} K{K} N{N} nv{nv}", out, ref)


def mlp_case(M, C):
    x = torch.randn(M, C)
    W1, b1 = torch.randn(4 * C, C) * C ** -0.5, torch.randn(4 * C) * 0.1
    W2, b2 = torch.randn(C, 4 * C) * (4 * C) ** -0.5, torch.randn(C) * 0.1
    lw, lb = torch.ones(C), torch.zeros(C)
    h = torch.nn.functional.layer_norm(x, (C,)) @ W1.t() + b1
    ref = x + torch.nn.functional.gelu(h) @ W2.t() + b2
    HC, TR = ops.mlp_config(C)
    Wp, b2p = packing.pack_mlp(W1.to(DEV), W2.to(DEV), b2.to(DEV), HC, TR)
    out = torch.zeros(M, C, device=DEV)
    ops.mlp(x.to(DEV), out, M, C, lw.to(DEV), lb.to(DEV), Wp, b1.to(DEV), b2p)
    torch.cuda.synchronize()
    report(f"mlp M{M} C{C} HC{HC} TR{TR}", out, ref)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
    if which == "gemm":
        for (M, K, N, nv) in [(128, 16, 16, None), (128, 64, 64, None), (128, 128, 64, None), (128, 64, 256, None),
                              (128, 64, 128, 64), (256, 192, 576, None), (128, 384, 1152, None), (128, 768, 384, 128)]:
            gemm_case(M, K, N, nv)
    else:
        for (M, C) in [(128, 16), (128, 48), (128, 64), (256, 96), (128, 192), (128, 384)]:
            mlp_case(M, C)
