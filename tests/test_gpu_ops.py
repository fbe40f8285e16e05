"""GPU parity tests, op level: every C-ABI entry point against the oracle's fp32 math on the same seeded
inputs.  Tolerances: bf16 tensor-core operands with fp32 accumulation -> max abs error <= 2e-2 of the
reference's max magnitude (north_star tolerance); fp32 CUDA-core kernels -> 1e-4."""
import math

import pytest
import torch

import swinwnet_b200 as S
from swinwnet_b200 import ops, packing
from oracle import swinwnet_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_BF16 = 2e-2
TOL_F32 = 1e-4
TOL_HEAD = 2e-3   # conv heads: TF32 tensor-core operands (10-bit mantissa), fp32 accumulation
OPD = S.ops.operand_dtype()   # 16-bit tensor-core operand dtype of the built library (fp16 default)


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.isfinite(a).all(), "non-finite values in kernel output"
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


def bf(x):  # bf16 rounding of an operand, as the kernel sees it
    return x.to(OPD).float()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,M", [(48, 300), (96, 128), (192, 257), (384, 130), (24, 1000), (12, 515)])
def test_rowgemm_ln_qkv(C, M):
    x = rnd(M, C, seed=1) * 2 + 0.3
    W, b = rnd(3 * C, C, seed=2, scale=C ** -0.5), rnd(3 * C, seed=3, scale=0.1)
    lw, lb = 1 + 0.1 * rnd(C, seed=4), 0.1 * rnd(C, seed=5)
    ref = O.linear(O.layer_norm(x, lw, lb), W, b)
    nv = packing.choose_chunk(3 * C, 256)
    Wp, bp, NT, nch = packing.pack_rowgemm(W.to(DEV), b.to(DEV), nv)
    out = torch.empty(M, 3 * C, device=DEV, dtype=OPD)
    ops.rowgemm(A=x.to(DEV), a_mode=ops.A_F32_LN, M=M, K=C, lda=C, ln_w=lw.to(DEV), ln_b=lb.to(DEV), Wp=Wp, NT=NT,
                nchunks=nch, n_valid=nv, e_mode=ops.E_BF16, bias=bp, out=out, ldo=3 * C)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_BF16


@pytest.mark.parametrize("C,M", [(48, 300), (192, 129), (384, 256), (12, 77)])
def test_rowgemm_bf16_proj_residual_alpha(C, M):
    a = rnd(M, C, seed=1).to(OPD)
    W, b = rnd(C, C, seed=2, scale=C ** -0.5), rnd(C, seed=3, scale=0.1)
    res = rnd(M, C, seed=4)
    gamma = torch.tensor([0.37])
    ref = res + gamma * O.linear(a.float(), W, b)
    nv = packing.choose_chunk(C, 256)
    Wp, bp, NT, nch = packing.pack_rowgemm(W.to(DEV), b.to(DEV), nv)
    out = torch.empty(M, C, device=DEV)
    ops.rowgemm(A=a.to(DEV), a_mode=ops.A_BF16, M=M, K=C, lda=C, Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_F32,
                bias=bp, out=out, ldo=C, res=res.to(DEV), ldres=C, alpha=gamma.to(DEV))
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_BF16


def test_rowgemm_inplace_residual():
    C, M = 96, 200
    a = rnd(M, C, seed=1).to(OPD)
    W, b = rnd(C, C, seed=2, scale=C ** -0.5), rnd(C, seed=3, scale=0.1)
    x = rnd(M, C, seed=4)
    ref = x + O.linear(a.float(), W, b)
    Wp, bp, NT, nch = packing.pack_rowgemm(W.to(DEV), b.to(DEV), C)
    xd = x.to(DEV)
    ops.rowgemm(A=a.to(DEV), a_mode=ops.A_BF16, M=M, K=C, lda=C, Wp=Wp, NT=NT, nchunks=nch, n_valid=C, e_mode=ops.E_F32,
                bias=bp, out=xd, ldo=C, res=xd, ldres=C)
    torch.cuda.synchronize()
    assert relerr(xd, ref) <= TOL_BF16


@pytest.mark.parametrize("C,H,W", [(48, 9, 13), (96, 8, 6), (192, 5, 5)])
def test_rowgemm_patch_merging(C, H, W):
    B = 2
    x = rnd(B, H * W, C, seed=1)
    sd = {"reduction.weight": rnd(2 * C, 4 * C, seed=2, scale=(4 * C) ** -0.5), "norm.weight": 1 + 0.1 * rnd(4 * C, seed=3),
          "norm.bias": 0.1 * rnd(4 * C, seed=4)}
    ref, res = O.patch_merging(sd, "", x, (H, W))
    m = S.model.PatchMerging(C).to(DEV)
    m.load_state_dict(sd)
    with torch.no_grad():
        out, res2 = m(x.to(DEV), (H, W))
    torch.cuda.synchronize()
    assert tuple(res2) == tuple(res)
    assert relerr(out, ref) <= TOL_BF16


@pytest.mark.parametrize("C,H,W,crop", [(384, 4, 6, None), (96, 5, 7, (9, 13)), (48, 6, 5, None), (24, 10, 12, None)])
def test_rowgemm_patch_expanding(C, H, W, crop):
    B = 2
    x = rnd(B, H * W, C, seed=1)
    sd = {"expand.weight": rnd(2 * C, C, seed=2, scale=C ** -0.5), "norm.weight": 1 + 0.1 * rnd(C // 2, seed=3),
          "norm.bias": 0.1 * rnd(C // 2, seed=4)}
    ref, res = O.patch_expanding(sd, "", x, (H, W))
    if crop is not None:
        ref = O.crop_tokens(ref, res, crop)
    m = S.model.PatchExpanding(C).to(DEV)
    m.load_state_dict(sd)
    with torch.no_grad():
        out, _ = m.run(x.to(DEV), (H, W), crop)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_BF16


@pytest.mark.parametrize("C,M", [(48, 300), (96, 128), (192, 200), (384, 129), (24, 700), (12, 1025)])
def test_fused_mlp(C, M):
    x = rnd(M, C, seed=1)
    W1, b1 = rnd(4 * C, C, seed=2, scale=C ** -0.5), rnd(4 * C, seed=3, scale=0.2)
    W2, b2 = rnd(C, 4 * C, seed=4, scale=(4 * C) ** -0.5), rnd(C, seed=5, scale=0.1)
    lw, lb = 1 + 0.1 * rnd(C, seed=6), 0.1 * rnd(C, seed=7)
    ref = x + O.linear(O.gelu_erf(O.linear(O.layer_norm(x, lw, lb), W1, b1)), W2, b2)
    HC, TR = ops.mlp_config(C)
    Wp, b2p = packing.pack_mlp(W1.to(DEV), W2.to(DEV), b2.to(DEV), HC, TR)
    xd = x.to(DEV)
    out = torch.empty_like(xd)
    ops.mlp(xd, out, M, C, lw.to(DEV), lb.to(DEV), Wp, b1.to(DEV), b2p)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_BF16
    ops.mlp(xd, xd, M, C, lw.to(DEV), lb.to(DEV), Wp, b1.to(DEV), b2p)  # in place
    torch.cuda.synchronize()
    assert relerr(xd, ref) <= TOL_BF16


def _win_attn_ref(qkv, bias, table, B, H, W, C, nH, shift):
    """oracle window attention fed with a given token-ordered qkv tensor (pads take the bias)."""
    ws, hd = 5, C // nH
    g = qkv.view(B, H, W, 3 * C)
    if shift:
        g = torch.roll(g, (-shift, -shift), (1, 2))
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    gp = bias.view(1, 1, 1, -1).expand(B, Hp, Wp, 3 * C).clone()
    gp[:, :H, :W] = g
    nWy, nWx = Hp // ws, Wp // ws
    win = gp.view(B, nWy, ws, nWx, ws, 3, nH, hd).permute(0, 1, 3, 5, 6, 2, 4, 7).reshape(B * nWy * nWx, 3, nH, 25, hd)
    q, k, v = win[:, 0] * hd ** -0.5, win[:, 1], win[:, 2]
    att = q @ k.transpose(-1, -2) + table[O.rel_pos_index(ws).reshape(-1)].view(25, 25, nH).permute(2, 0, 1)[None]
    if shift:
        rid = O.shift_region_ids(Hp, Wp, ws, shift).view(nWy, ws, nWx, ws).permute(0, 2, 1, 3).reshape(nWy * nWx, 25)
        m = torch.where(rid[:, :, None] == rid[:, None, :], 0.0, -100.0)
        att = (att.view(B, nWy * nWx, nH, 25, 25) + m[None, :, None]).view(-1, nH, 25, 25)
    o = (torch.softmax(att, -1) @ v).permute(0, 2, 1, 3).reshape(B, nWy, nWx, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    o = o.reshape(B, Hp, Wp, C)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    return o[:, :H, :W].reshape(B * H * W, C)


@pytest.mark.parametrize("C,nH,H,W,shift", [(48, 3, 10, 15, 0), (96, 6, 7, 11, 0), (384, 24, 4, 6, 0), (384, 12, 8, 5, 0),
                                            (192, 6, 5, 10, 0), (24, 3, 10, 10, 0), (12, 3, 12, 9, 0), (48, 3, 10, 15, 2),
                                            (96, 3, 5, 5, 3),
                                            # warp-per-window kernel (shift 0, head_dim 16 / 32): padded windows in y and x, an image
                                            # smaller than a window, many windows per warp
                                            (192, 6, 63, 120, 0), (192, 12, 32, 60, 0), (384, 12, 32, 61, 0), (384, 24, 16, 30, 0),
                                            (96, 3, 3, 4, 0), (96, 6, 125, 240, 0)])
def test_window_attention(C, nH, H, W, shift):
    B = 2
    qkv = rnd(B * H * W, 3 * C, seed=1).to(OPD)
    bias, table = rnd(3 * C, seed=2, scale=0.3), rnd(81, nH, seed=3, scale=0.5)
    ref = _win_attn_ref(qkv.float(), bf(bias), table, B, H, W, C, nH, shift)
    out = torch.empty(B * H * W, C, device=DEV, dtype=OPD)
    out.fill_(float("nan"))
    ops.window_attention(qkv.to(DEV), out, bias.to(DEV), table.to(DEV), B, H, W, C, nH, shift)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= 1e-2
    if shift == 0:      # bias images passed in (what the model does) == built in the kernel from the table
        frags = packing.rel_pos_bias_fragments(table.to(DEV), 1.4426950408889634).reshape(-1).contiguous()
        out2 = torch.full_like(out, float("nan"))
        ops.window_attention(qkv.to(DEV), out2, bias.to(DEV), table.to(DEV), B, H, W, C, nH, 0, frags)
        torch.cuda.synchronize()
        assert torch.equal(out2, out) or relerr(out2, ref) <= 1e-2


@pytest.mark.parametrize("C,Lq,Lk", [(192, 100, 75), (384, 70, 130), (192, 64, 64), (384, 12, 40), (192, 480, 1920)])
def test_cross_attention_core(C, Lq, Lk):
    B, nH = 2, 3
    hd = C // nH
    q = rnd(B, Lq, C, seed=1).to(OPD)
    kv = rnd(B, Lk, 2 * C, seed=2).to(OPD)
    Q = q.float().view(B, Lq, nH, hd).permute(0, 2, 1, 3)
    K = kv.float()[..., :C].reshape(B, Lk, nH, hd).permute(0, 2, 1, 3)
    V = kv.float()[..., C:].reshape(B, Lk, nH, hd).permute(0, 2, 1, 3)
    ref = (torch.softmax(Q @ K.transpose(-1, -2) * hd ** -0.5, -1) @ V).permute(0, 2, 1, 3).reshape(B, Lq, C)
    out = torch.empty(B, Lq, C, device=DEV, dtype=OPD)
    ops.cross_attention(q.to(DEV), kv.to(DEV), out, B, Lq, Lk, C, nH)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= 1.5e-2


@pytest.mark.parametrize("Cin,H,W,s", [(2, 40, 60, 1), (1, 35, 51, 1), (2, 80, 120, 2), (1, 250, 480, 1)])
def test_patch_embed(Cin, H, W, s):
    B = 2
    x = rnd(B, Cin, H, W, seed=1) * 3
    sd = {"proj.weight": rnd(48, Cin, 2, 2, seed=2, scale=0.5), "proj.bias": rnd(48, seed=3, scale=0.1),
          "norm.weight": 1 + 0.1 * rnd(48, seed=4), "norm.bias": 0.1 * rnd(48, seed=5)}
    ref, pres = O.patch_embed(sd, "", x, s)
    m = S.model.ScaleAwarePatchEmbed(2, Cin, 48).to(DEV)
    m.load_state_dict(sd)
    with torch.no_grad():
        out, pres2 = m(x.to(DEV), scale_factor=s)
    torch.cuda.synchronize()
    assert tuple(pres) == tuple(pres2)
    assert relerr(out, ref) <= TOL_F32


@pytest.mark.parametrize("scale", [1, 2])
def test_segmentation_head(scale):
    B, Hq, Wq = 2, 9, 14
    x = rnd(B, Hq * Wq, 48, seed=1)
    sd = {"seg_head.0.weight": rnd(24, 48, 3, 3, seed=2, scale=0.05), "seg_head.0.bias": rnd(24, seed=3, scale=0.1),
          "seg_head.2.weight": rnd(1, 24, 1, 1, seed=4, scale=0.2), "seg_head.2.bias": rnd(1, seed=5, scale=0.1)}
    res = (Hq * 2 * scale, Wq * 2 * scale)
    ref = O.segmentation_head(sd, "", x, res, scale)
    m = S.model.SegmentationHead().to(DEV)
    m.load_state_dict(sd)
    with torch.no_grad():
        out = m(x.to(DEV), res, scale_factor=scale)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_HEAD


def test_recon_head_and_glue():
    B, H, W = 2, 12, 20
    x = rnd(B, H * W, 12, seed=1)
    w1, b1 = rnd(12, 12, 3, 3, seed=2, scale=0.1), rnd(12, seed=3, scale=0.1)
    w2, b2 = rnd(2, 12, 1, 1, seed=4, scale=0.3), rnd(2, seed=5, scale=0.1)
    h = O.gelu_erf(O.conv3x3_nhwc(x.view(B, H, W, 12), w1, b1))
    ref = (h @ w2.view(2, 12).t() + b2).permute(0, 3, 1, 2)[:, :, :10, :18]
    out = torch.empty(B, 2, 10, 18, device=DEV)
    ops.recon_head(x.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), out, B, H, W, 2, 10, 18)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_HEAD
    full = (h @ w2.view(2, 12).t() + b2).permute(0, 3, 1, 2)            # uncropped, row length a multiple of 4
    out4 = torch.empty(B, 2, H, W, device=DEV)
    ops.recon_head(x.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), out4, B, H, W, 2, H, W)
    torch.cuda.synchronize()
    assert relerr(out4, full) <= TOL_HEAD
    # glue: ensure_2ch + sigmoid mask + minmax + normalize / denormalize round trip
    img = rnd(B, 1, H, W, seed=6).abs() * 100 + 5
    seg = rnd(B, 1, H, W, seed=7)
    images, seg_map, masked, mm = ops.sigmoid_mask(img.to(DEV), seg.to(DEV), ensure_2ch=True, want_minmax=True)
    ref_img = O.ensure_2ch(img)
    ref_masked = ref_img * torch.sigmoid(seg)
    ref_norm, params = O.normalize_piecewise(ref_masked)
    assert relerr(images, ref_img) <= 1e-6 and relerr(seg_map, torch.sigmoid(seg)) <= 1e-5
    assert relerr(masked, ref_masked) <= 1e-5
    assert relerr(mm.view(B, 2, 2)[..., 0], params[0].view(B, 2)) <= 1e-5
    assert relerr(mm.view(B, 2, 2)[..., 1], params[1].view(B, 2)) <= 1e-5
    norm = ops.normalize(masked, mm, inverse=False)
    assert relerr(norm, ref_norm) <= 1e-5
    den = ops.normalize(norm, mm, inverse=True)
    assert relerr(den, O.denormalize_piecewise(ref_norm, params)) <= 1e-5


@pytest.mark.parametrize("C,H,W,shift", [(12, 10, 15, 0), (24, 7, 11, 0), (12, 13, 9, 0), (24, 10, 10, 2), (12, 5, 5, 3)])
def test_swin_block_small_fused(C, H, W, shift):
    """whole-block fp32 kernel for the UpscalingHead widths against the oracle block (incl. padding and shift)."""
    B, nH = 2, 3
    x = rnd(B, H * W, C, seed=1)
    shapes = {"norm1.weight": (C,), "norm1.bias": (C,), "attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
              "attn.relative_position_bias_table": (81, nH), "attn.proj.weight": (C, C), "attn.proj.bias": (C,),
              "norm2.weight": (C,), "norm2.bias": (C,), "mlp.0.weight": (4 * C, C), "mlp.0.bias": (4 * C,),
              "mlp.3.weight": (C, 4 * C), "mlp.3.bias": (C,)}
    sd = {k: rnd(*s, seed=10 + i) * (0.3 if "weight" in k and len(s) == 2 else 0.2) for i, (k, s) in enumerate(shapes.items())}
    sd["norm1.weight"] += 1.0
    sd["norm2.weight"] += 1.0
    ref = O.swin_block(sd, "", x, (H, W), nH, shift)
    params = [sd[k].to(DEV).contiguous() for k in shapes]
    xd = x.to(DEV)
    out = torch.empty_like(xd)
    ops.swin_block_small(xd, out, B, H, W, C, nH, shift, 1e-5, params)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= 2e-5
    ops.swin_block_small(xd, xd, B, H, W, C, nH, shift, 1e-5, params)   # in place
    torch.cuda.synchronize()
    assert relerr(xd, ref) <= 2e-5


def _block_sd(C, nH):
    shapes = {"norm1.weight": (C,), "norm1.bias": (C,), "attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
              "attn.relative_position_bias_table": (81, nH), "attn.proj.weight": (C, C), "attn.proj.bias": (C,),
              "norm2.weight": (C,), "norm2.bias": (C,), "mlp.0.weight": (4 * C, C), "mlp.0.bias": (4 * C,),
              "mlp.3.weight": (C, 4 * C), "mlp.3.bias": (C,)}
    sd = {k: rnd(*s, seed=10 + i) * ((s[-1] ** -0.5) if "weight" in k and len(s) == 2 else 0.2)
          for i, (k, s) in enumerate(shapes.items())}
    sd["norm1.weight"] += 1.0
    sd["norm2.weight"] += 1.0
    return sd, list(shapes)


@pytest.mark.parametrize("C,nH,B,H,W,do_mlp", [
    (48, 3, 2, 10, 15, True), (48, 3, 1, 13, 9, True), (48, 3, 3, 25, 40, True), (48, 3, 2, 7, 11, False),
    (24, 3, 2, 10, 15, True), (24, 3, 1, 32, 61, True), (12, 3, 2, 13, 9, True), (12, 3, 1, 40, 65, True),
    (12, 3, 2, 10, 10, False), (48, 6, 1, 12, 23, True),
    (48, 3, 2, 100, 101, True), (24, 3, 1, 180, 175, True), (12, 3, 1, 240, 245, True)])   # several tiles per persistent CTA
def test_swin_block_fused(C, nH, B, H, W, do_mlp):
    """tcgen05 whole-block kernel (csrc/swin_fused.cu) against the oracle block: window padding (H, W not multiples of
    5), ragged last tile (window count not a multiple of 5), several tiles per CTA, attention-only mode."""
    x = rnd(B, H * W, C, seed=1) * 1.5 + 0.2
    sd, order = _block_sd(C, nH)
    if do_mlp:
        ref = O.swin_block(sd, "", x, (H, W), nH, 0)
    else:
        ref = x + O.window_attention(sd, "attn.", O.layer_norm(x, sd["norm1.weight"], sd["norm1.bias"]), (H, W), nH, 0)
    d = {k: v.to(DEV) for k, v in sd.items()}
    Wpk, fpk = packing.pack_fused_block(d["norm1.weight"], d["norm1.bias"], d["attn.qkv.weight"], d["attn.qkv.bias"],
                                        d["attn.relative_position_bias_table"], d["attn.proj.weight"], d["attn.proj.bias"],
                                        d["norm2.weight"], d["norm2.bias"], d["mlp.0.weight"], d["mlp.0.bias"],
                                        d["mlp.3.weight"], d["mlp.3.bias"], nH)
    xd = x.to(DEV)
    out = torch.full_like(xd, float("nan"))
    ops.swin_block_fused(xd, out, B, H, W, C, nH, 1e-5, Wpk, fpk, do_mlp)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_BF16 / 4


@pytest.mark.parametrize("C,B,H,W", [
    (12, 2, 10, 15), (12, 2, 13, 9), (12, 1, 40, 65), (12, 1, 5, 5), (12, 1, 3, 4), (12, 1, 240, 245), (12, 3, 100, 190),
    (24, 2, 10, 15), (24, 1, 32, 61), (24, 1, 5, 5), (24, 1, 2, 7), (24, 1, 180, 175), (24, 2, 125, 240),
    (48, 2, 10, 15), (48, 1, 13, 9), (48, 1, 3, 4), (48, 3, 25, 40), (48, 2, 100, 101), (48, 2, 125, 240),
    (48.6, 1, 12, 23), (48.6, 2, 63, 121)])
def test_swin_block_warp(C, B, H, W):
    """one-warp-per-window block kernel (csrc/swin_warp.cu) against the oracle block: window padding (H, W not multiples
    of 5: zero tokens AFTER norm1 that still act as keys), images smaller than a window, many windows per warp;
    C = 48.6 stands for 48 channels with 6 heads"""
    nH = 6 if C == 48.6 else 3
    C = int(C)
    x = rnd(B, H * W, C, seed=1) * 1.5 + 0.2
    sd, order = _block_sd(C, nH)
    ref = O.swin_block(sd, "", x, (H, W), nH, 0)
    d = {k: v.to(DEV) for k, v in sd.items()}
    Wpk, fpk = packing.pack_warp_block(*[d[k] for k in order], nH)
    xd = x.to(DEV)
    out = torch.full_like(xd, float("nan"))
    ops.swin_block_warp(xd, out, B, H, W, C, nH, 1e-5, Wpk, fpk)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_BF16 / 4
    out2 = torch.full_like(xd, float("nan"))
    ops.swin_block_warp(xd, out2, B, H, W, C, nH, 1e-5, Wpk, fpk)
    assert torch.equal(out, out2)                                   # deterministic
    with pytest.raises(RuntimeError, match="alias"):
        ops.swin_block_warp(xd, xd, B, H, W, C, nH, 1e-5, Wpk, fpk)


@pytest.mark.parametrize("C,B,H,W,depth", [(12, 2, 13, 9, 2), (12, 1, 100, 190, 2), (24, 1, 32, 61, 2), (24, 2, 125, 240, 3),
                                           (12, 1, 40, 65, 4), (48, 2, 13, 9, 2), (48, 2, 125, 240, 2), (48, 1, 32, 61, 3)])
def test_swin_block_warp_layer(C, B, H, W, depth):
    """`depth` consecutive blocks (different weights) in one launch == the oracle blocks applied one after the other"""
    nH = 3
    x = rnd(B, H * W, C, seed=2) * 1.5 + 0.2
    ref, Wl, fl = x, [], []
    for i in range(depth):
        sd, order = _block_sd(C, nH)
        sd = {k: v * (1.0 + 0.05 * i) + (0.01 * i if k.endswith("bias") else 0.0) for k, v in sd.items()}
        ref = O.swin_block(sd, "", ref, (H, W), nH, 0)
        Wp, fp = packing.pack_warp_block(*[sd[k].to(DEV) for k in order], nH)
        Wl.append(Wp)
        fl.append(fp)
    xd = x.to(DEV)
    out = torch.full_like(xd, float("nan"))
    ops.swin_block_warp(xd, out, B, H, W, C, nH, 1e-5, torch.cat(Wl), torch.cat(fl), depth)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_BF16 / 4 * depth ** 0.5
    with pytest.raises(RuntimeError, match="depth" if C < 48 else "depth|shared memory"):
        ops.swin_block_warp(xd, out, B, H, W, C, nH, 1e-5, torch.cat(Wl + Wl)[:5 * Wl[0].numel()], torch.cat(fl + fl)[:5 * fl[0].numel()], 5)
    if C == 48 and depth == 3:      # four blocks of C = 48 do not fit in shared memory: loud error, nothing launched
        with pytest.raises(RuntimeError, match="shared memory"):
            ops.swin_block_warp(xd, out, B, H, W, C, nH, 1e-5, torch.cat(Wl + Wl[:1]), torch.cat(fl + fl[:1]), 4)


@pytest.mark.parametrize("nH,B,H,W", [(3, 2, 10, 15), (6, 1, 13, 9), (3, 1, 63, 120), (6, 2, 100, 101)])
def test_swin_attn_stream_c96(nH, B, H, W):
    """streamed-weight C=96 attention-half kernel: x + proj(W-MSA(LN1 x)) against the oracle (padding, ragged last
    tile, several tiles per persistent CTA so that the weight ring wraps)."""
    C = 96
    x = rnd(B, H * W, C, seed=1) * 1.5 + 0.2
    sd, _ = _block_sd(C, nH)
    ref = x + O.window_attention(sd, "attn.", O.layer_norm(x, sd["norm1.weight"], sd["norm1.bias"]), (H, W), nH, 0)
    d = {k: v.to(DEV) for k, v in sd.items()}
    Wpk, fpk = packing.pack_fused_attn_stream(d["norm1.weight"], d["norm1.bias"], d["attn.qkv.weight"], d["attn.qkv.bias"],
                                              d["attn.relative_position_bias_table"], d["attn.proj.weight"],
                                              d["attn.proj.bias"], nH)
    xd = x.to(DEV)
    out = torch.full_like(xd, float("nan"))
    ops.swin_block_fused(xd, out, B, H, W, C, nH, 1e-5, Wpk, fpk, False)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL_BF16 / 4


def test_dspace_histogram_matches_oracle():
    """physics front end: batched I(d) histogram kernel against the oracle's restatement of
    Qwrapper.tensor_to_d, HR and LR geometries of the reference protocol (tests.py:168-172)."""
    import numpy as np
    from oracle import diffraction_metrics_oracle as DM
    from swinwnet_b200 import physics
    for (H, W, centers) in ((500, 960, DM.D_CENTERS_HR), (250, 480, DM.D_CENTERS_LR)):
        x = rnd(3, 2, H, W, seed=H).abs() * 100 + 1
        q = physics.Qwrapper(fixed_centers=centers, device=DEV)
        res = q.tensor_to_d(x.to(DEV))
        assert len(res) == 3 and res[0]["d"].shape == (len(centers),)
        for b in range(3):
            d_ref, I_ref = DM.to_d_space(x[b, 0].numpy(), centers)
            assert np.allclose(res[b]["d"], d_ref)
            assert np.abs(res[b]["I"] - I_ref).max() <= 1e-5 * np.abs(I_ref).max()


def test_copy_cols():
    src = rnd(37, 48, seed=1).to(DEV)
    dst = torch.zeros(37, 96, device=DEV)
    ops.copy_cols(src, 48, dst, 48, 96, 37, 48)
    torch.cuda.synchronize()
    assert torch.equal(dst[:, 48:], src) and dst[:, :48].abs().max().item() == 0


def test_errors_are_loud():
    with pytest.raises(RuntimeError):
        ops.rowgemm(A=torch.zeros(4, 6, device=DEV), a_mode=ops.A_F32, M=4, K=6, lda=6, Wp=torch.zeros(8, device=DEV),
                    NT=16, nchunks=1, n_valid=4, e_mode=ops.E_F32, out=torch.zeros(4, 4, device=DEV), ldo=4)
    with pytest.raises(RuntimeError):
        ops.window_attention(torch.zeros(25, 30, device=DEV, dtype=OPD),
                             torch.zeros(25, 10, device=DEV, dtype=OPD), torch.zeros(30, device=DEV),
                             torch.zeros(81, 1, device=DEV), 1, 5, 5, 10, 1, 0)   # head_dim 10 unsupported
    with pytest.raises(RuntimeError):
        S.SwinWNet(depths=[2, 2, 2, 2]).segment_1(torch.zeros(1, 1, 20, 20))    # CPU tensor: no fallback
