"""Target of the compute-sanitizer runs: every kernel family once, at shapes that give the persistent
kernels several tiles per CTA and wrap the weight / accumulator rings, plus one small ST-pipeline pass.  Sizes are small:
under the sanitizer kernels run 10-100x slower.   compute-sanitizer --tool memcheck|racecheck python tools/sanitize_target.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import benchdata  # noqa: E402
import swinwnet_b200 as S  # noqa: E402
from swinwnet_b200 import ops, packing  # noqa: E402

DEV = "cuda"
torch.manual_seed(0)


def block_params(C, nH):
    shp = [(C,), (C,), (3 * C, C), (3 * C,), (81, nH), (C, C), (C,), (C,), (C,), (4 * C, C), (4 * C,), (C, 4 * C), (C,)]
    return [torch.randn(*s, device=DEV) * 0.1 for s in shp]


# fused whole-block kernels (C = 12 / 24 / 48) and the streamed C = 96 W-MSA kernel: ~3 tiles per CTA
for C, nH, B, H, W in ((12, 3, 2, 250, 240), (24, 3, 2, 125, 240), (48, 3, 4, 125, 120), (96, 3, 4, 125, 120), (96, 6, 8, 63, 120)):
    x = torch.randn(B, H * W, C, device=DEV)
    out = torch.empty_like(x)
    p = block_params(C, nH)
    if C < 96:
        Wpk, fpk = packing.pack_fused_block(*p, nH)
    else:
        Wpk, fpk = packing.pack_fused_attn_stream(*p[:7], nH)
    ops.swin_block_fused(x, out, B, H, W, C, nH, 1e-5, Wpk, fpk, C < 96)
    torch.cuda.synchronize()
    print("fused", C, nH, float(out.abs().mean()))

# MLP kernels: persistent (C <= 96, 3 tiles per CTA) and per-tile (C = 192 / 384)
for C, M in ((48, 128 * 148 * 3 + 5), (96, 128 * 148 * 3 + 5), (192, 128 * 200 + 7), (384, 128 * 150 + 1)):
    x = torch.randn(M, C, device=DEV)
    W1, b1 = torch.randn(4 * C, C, device=DEV) * C ** -0.5, torch.zeros(4 * C, device=DEV)
    W2, b2 = torch.randn(C, 4 * C, device=DEV) * (4 * C) ** -0.5, torch.zeros(C, device=DEV)
    HC, TR = ops.mlp_config(C)
    Wp, b2p = packing.pack_mlp(W1, W2, b2, HC, TR)
    out = torch.empty_like(x)
    ops.mlp(x, out, M, C, torch.ones(C, device=DEV), torch.zeros(C, device=DEV), Wp, b1, b2p)
    torch.cuda.synchronize()
    print("mlp", C, float(out.abs().mean()))

# row GEMMs: persistent (K = 96) and per-tile with LayerNorm prologue (K = 192, 384)
for K, N, ln, M in ((96, 48, False, 128 * 148 * 3 + 9), (192, 576, True, 128 * 160 + 3), (384, 1152, True, 128 * 150)):
    A = torch.randn(M, K, device=DEV)
    Wq, bq = torch.randn(N, K, device=DEV) * K ** -0.5, torch.zeros(N, device=DEV)
    nv = packing.choose_chunk(N, 128 if K <= 192 else 256)
    Wp, bp, NT, nch = packing.pack_rowgemm(Wq, bq, nv)
    out = torch.empty(M, N, device=DEV, dtype=ops.operand_dtype() if ln else torch.float32)
    ops.rowgemm(A=A, a_mode=ops.A_F32_LN if ln else ops.A_F32, M=M, K=K, lda=K, ln_w=torch.ones(K, device=DEV), ln_b=torch.zeros(K, device=DEV),
                Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_BF16 if ln else ops.E_F32, bias=bp, out=out, ldo=N)
    torch.cuda.synchronize()
    print("rowgemm", K, N, float(out.float().abs().mean()))

# one ST-pipeline pass (all remaining kernels: patch embed, merge / expand, window + cross attention, heads, glue)
man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
m = S.SwinWNet(error_matrix=True, depths=[2, 2, 2, 2])
m.load_state_dict(benchdata.make_state_dict(man["wnet_em"], seed=1), strict=True)
inf = S.SwinWNetInference(m, DEV)
y = inf(benchdata.synthetic_diffractions(2, seed=3, H=60, W=80, two_channel=False).to(DEV))
torch.cuda.synchronize()
print("pipeline", tuple(y.shape), float(y.abs().mean()))
print("SANITIZE TARGET DONE")
