"""CPU/fp32 ORACLE for the SwinWNet forward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch functional restatement (plain torch fp32 tensor math, no
nn.Module, explicit index formulas) of the algorithm in the reference's
``SwinWNet.py`` / ``ST_Inference_Pipline.py``.  It is the checker for the CUDA
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
never imports it and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no golden vectors / unit tests for this path
(SURVEY.md §8c), so the oracle is pinned against the *reference itself executed in
the authoring container*: ``oracle/make_golden.py`` imports the unmodified
``/root/reference/SwinWNet.py`` and stores seeded input/output vectors under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through this
file (max abs err <= 2e-5 on every stage tensor).

All functions take a flat ``state_dict``-style mapping ``sd`` whose keys are the
reference's parameter names and a key ``prefix``.
File:line citations refer to /root/reference/.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]

WS = 5  # window size used by every shipped model (SwinWNet.py:802)


# ----------------------------------------------------------------------------
# elementary pieces
# ----------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance (SwinWNet.py:220,226,287,395)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (SwinWNet.py:230,503,645)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def linear(x: Tensor, w: Tensor, b: Tensor | None = None) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


def rel_pos_index(ws: int = WS) -> Tensor:
    """index[i,j] = (yi-yj+ws-1)*(2ws-1) + (xi-xj+ws-1)  (SwinWNet.py:163-172)."""
    t = torch.arange(ws * ws)
    y, x = t // ws, t % ws
    return (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


def shift_region_ids(Hp: int, Wp: int, ws: int, shift: int) -> Tensor:
    """Region id (0..8) per padded-grid cell for the shifted-window mask.

    Follows the slice construction at SwinWNet.py:133-140.  NOTE: the reference then
    builds the mask with a shape bug (SwinWNet.py:147) and never uses shift>0
    (SwinWNet.py:328); the oracle implements the standard-Swin semantics
    mask[w,i,j] = 0 if region(i)==region(j) else -100 — an
    extension beyond what the reference can execute.
    """
    def band(n):
        ids = torch.zeros(n, dtype=torch.long)
        ids[n - ws: n - shift] = 1
        ids[n - shift:] = 2
        return ids
    return band(Hp)[:, None] * 3 + band(Wp)[None, :]


# ----------------------------------------------------------------------------
# patch embed  (SwinWNet.py:53-82)
# ----------------------------------------------------------------------------
def patch_embed(sd: SD, prefix: str, x: Tensor, scale: int = 1, patch: int = 2):
    B, Cin, H, W = x.shape
    ps = patch
    # reference pad formula with its literal operator precedence (SwinWNet.py:70-71)
    pad_h = (ps * scale - H % ps * scale) % ps * scale
    pad_w = (ps * scale - W % ps * scale) % ps * scale
    if pad_h or pad_w:
        x = torch.nn.functional.pad(x, (0, pad_w, 0, pad_h))
    Hn, Wn = H + pad_h, W + pad_w
    w = sd[prefix + "proj.weight"]  # [E, Cin, ps, ps]
    bias = sd[prefix + "proj.bias"]
    stride = ps * scale
    Ho = (Hn - scale * (ps - 1) - 1) // stride + 1
    Wo = (Wn - scale * (ps - 1) - 1) // stride + 1
    acc = bias.view(1, 1, 1, -1).expand(B, Ho, Wo, -1).clone()
    for a in range(ps):
        for b in range(ps):
            tap = x[:, :, a * scale: a * scale + (Ho - 1) * stride + 1: stride,
                    b * scale: b * scale + (Wo - 1) * stride + 1: stride]  # [B,Cin,Ho,Wo]
            acc = acc + torch.einsum("bchw,ec->bhwe", tap, w[:, :, a, b])
    t = acc.reshape(B, Ho * Wo, -1)
    t = layer_norm(t, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    return t, (Hn, Wn)


# ----------------------------------------------------------------------------
# Swin block  (SwinWNet.py:183-209, 236-280)
# ----------------------------------------------------------------------------
def window_attention(sd: SD, prefix: str, xn: Tensor, res: Tuple[int, int], num_heads: int,
                     shift: int = 0, ws: int = WS) -> Tensor:
    """xn: post-norm1 tokens [B, H*W, C] -> attention branch output [B, H*W, C]."""
    B, L, C = xn.shape
    H, W = res
    hd = C // num_heads
    g = xn.view(B, H, W, C)
    if shift > 0:
        g = torch.roll(g, shifts=(-shift, -shift), dims=(1, 2))
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    gp = torch.zeros(B, Hp, Wp, C, dtype=xn.dtype, device=xn.device)
    gp[:, :H, :W] = g                      # zero pad AFTER the norm (SwinWNet.py:242,254)
    nWy, nWx = Hp // ws, Wp // ws
    win = gp.view(B, nWy, ws, nWx, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B * nWy * nWx, ws * ws, C)
    qkv = linear(win, sd[prefix + "qkv.weight"], sd.get(prefix + "qkv.bias"))
    qkv = qkv.view(-1, ws * ws, 3, num_heads, hd)
    q = qkv[:, :, 0].permute(0, 2, 1, 3) * (hd ** -0.5)
    k = qkv[:, :, 1].permute(0, 2, 1, 3)
    v = qkv[:, :, 2].permute(0, 2, 1, 3)
    att = q @ k.transpose(-1, -2)                                   # [Bw,nH,25,25]
    table = sd[prefix + "relative_position_bias_table"]            # [81,nH]
    bias = table[rel_pos_index(ws).reshape(-1).to(table.device)].view(ws * ws, ws * ws, num_heads).permute(2, 0, 1)
    att = att + bias.unsqueeze(0)
    if shift > 0:
        rid = shift_region_ids(Hp, Wp, ws, shift).to(xn.device)
        rw = rid.view(nWy, ws, nWx, ws).permute(0, 2, 1, 3).reshape(nWy * nWx, ws * ws)
        m = torch.where(rw[:, :, None] == rw[:, None, :], 0.0, -100.0).to(xn.dtype)       # [nW,25,25]
        att = (att.view(B, nWy * nWx, num_heads, ws * ws, ws * ws) + m[None, :, None]).view(-1, num_heads, ws * ws, ws * ws)
    att = torch.softmax(att, dim=-1)
    o = (att @ v).permute(0, 2, 1, 3).reshape(-1, ws * ws, C)
    o = linear(o, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])
    o = o.view(B, nWy, nWx, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))   # on the padded grid, as SwinWNet.py:262-266
    return o[:, :H, :W].reshape(B, L, C)


def swin_block(sd: SD, prefix: str, x: Tensor, res, num_heads: int, shift: int = 0) -> Tensor:
    B, L, C = x.shape
    assert L == res[0] * res[1], "input feature has wrong size"
    xn = layer_norm(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"])
    x = x + window_attention(sd, prefix + "attn.", xn, res, num_heads, shift)
    h = layer_norm(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"])
    h = gelu_erf(linear(h, sd[prefix + "mlp.0.weight"], sd[prefix + "mlp.0.bias"]))
    return x + linear(h, sd[prefix + "mlp.3.weight"], sd[prefix + "mlp.3.bias"])


def basic_layer(sd: SD, prefix: str, x: Tensor, res, depth: int, num_heads: int) -> Tensor:
    """BasicLayer: every block has shift_size=0 (SwinWNet.py:328)."""
    for d in range(depth):
        x = swin_block(sd, f"{prefix}blocks.{d}.", x, res, num_heads, 0)
    return x


# ----------------------------------------------------------------------------
# merge / expand  (SwinWNet.py:289-316, 397-412)
# ----------------------------------------------------------------------------
def patch_merging(sd: SD, prefix: str, x: Tensor, res):
    B, L, C = x.shape
    H, W = res
    assert L == H * W, "input feature has wrong size"
    He, We = H + (H & 1), W + (W & 1)
    g = torch.zeros(B, He, We, C, dtype=x.dtype, device=x.device)
    g[:, :H, :W] = x.view(B, H, W, C)     # zero pad BEFORE the norm (SwinWNet.py:295-312)
    cat = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], -1)
    cat = cat.reshape(B, -1, 4 * C)
    cat = layer_norm(cat, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    return linear(cat, sd[prefix + "reduction.weight"]), (He // 2, We // 2)


def patch_expanding(sd: SD, prefix: str, x: Tensor, res):
    B, L, C = x.shape
    H, W = res
    assert L == H * W, "input feature has wrong size"
    e = linear(x, sd[prefix + "expand.weight"]).view(B, H, W, 2, 2, C // 2)
    out = torch.empty(B, 2 * H, 2 * W, C // 2, dtype=x.dtype, device=x.device)
    for i in range(2):
        for j in range(2):
            out[:, i::2, j::2] = e[:, :, :, i, j]      # pixel (2h+i,2w+j) <- channel group 2i+j
    out = out.reshape(B, 4 * H * W, C // 2)
    return layer_norm(out, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"]), (2 * H, 2 * W)


def crop_tokens(x: Tensor, cur, tgt) -> Tensor:
    B, L, C = x.shape
    assert cur[0] >= tgt[0] and cur[1] >= tgt[1]
    return x.view(B, cur[0], cur[1], C)[:, :tgt[0], :tgt[1]].reshape(B, tgt[0] * tgt[1], C)


# ----------------------------------------------------------------------------
# encoder / bottleneck / decoder  (SwinWNet.py:362-378, 387-388, 468-493)
# ----------------------------------------------------------------------------
def encoder(sd: SD, prefix: str, x: Tensor, res, depths: Sequence[int], heads: Sequence[int]):
    skips, rlist = [], []
    n = len(depths)
    for i in range(n - 1):
        x = basic_layer(sd, f"{prefix}layers.{i}.", x, res, depths[i], heads[i])
        skips.append(x)
        rlist.append(res)
        x, res = patch_merging(sd, f"{prefix}downs.{i}.", x, res)
    x = basic_layer(sd, f"{prefix}layers.{n - 1}.", x, res, depths[-1], heads[-1])
    skips.append(x)
    rlist.append(res)
    return skips, rlist, res


def bottleneck(sd: SD, prefix: str, x: Tensor, res, heads_last: int) -> Tensor:
    return basic_layer(sd, prefix + "layer.", x, res, 2, heads_last)


def decoder(sd: SD, prefix: str, x: Tensor, res, skips: List[Tensor], rlist, depths, heads):
    dskips, dres = skips[-2::-1], rlist[-2::-1]
    ddepths, dheads = list(depths)[-2::-1], list(heads)[-2::-1]
    for i in range(len(depths) - 1):
        x, nres = patch_expanding(sd, f"{prefix}ups.{i}.", x, res)
        if tuple(nres) != tuple(dres[i]):
            x = crop_tokens(x, nres, dres[i])
        x = torch.cat([x, dskips[i]], dim=-1)                # expanded first, skip second
        x = basic_layer(sd, f"{prefix}swin_blocks.{i}.", x, dres[i], ddepths[i], dheads[i])
        x = linear(x, sd[f"{prefix}linears.{i}.weight"], sd[f"{prefix}linears.{i}.bias"])
        res = dres[i]
    return x, res


# ----------------------------------------------------------------------------
# heads  (SwinWNet.py:507-531, 656-688)
# ----------------------------------------------------------------------------
def conv3x3_nhwc(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """x [B,H,W,Ci], w [Co,Ci,3,3], zero padding 1 -> [B,H,W,Co]."""
    B, H, W, Ci = x.shape
    xp = torch.zeros(B, H + 2, W + 2, Ci, dtype=x.dtype, device=x.device)
    xp[:, 1:-1, 1:-1] = x
    acc = b.view(1, 1, 1, -1).expand(B, H, W, -1).clone()
    for dy in range(3):
        for dx in range(3):
            acc = acc + xp[:, dy:dy + H, dx:dx + W] @ w[:, :, dy, dx].t()
    return acc


def bilinear_up(x: Tensor, s: int) -> Tensor:
    """F.interpolate(mode='bilinear', align_corners=False, scale_factor=s) on [B,h,w]."""
    B, h, w = x.shape

    def taps(n):
        d = torch.arange(n * s, dtype=torch.float32, device=x.device)
        src = torch.clamp((d + 0.5) / s - 0.5, min=0.0)
        i0 = src.floor().long().clamp(max=n - 1)
        i1 = (i0 + 1).clamp(max=n - 1)
        f = src - i0.float()
        return i0, i1, f

    y0, y1, fy = taps(h)
    x0, x1, fx = taps(w)
    top = x[:, y0][:, :, x0] * (1 - fx) + x[:, y0][:, :, x1] * fx
    bot = x[:, y1][:, :, x0] * (1 - fx) + x[:, y1][:, :, x1] * fx
    return top * (1 - fy)[None, :, None] + bot * fy[None, :, None]


def segmentation_head(sd: SD, prefix: str, x: Tensor, padded_res, scale: int = 1, patch: int = 2):
    B, N, C = x.shape
    H, W = padded_res
    Hq, Wq = H // (patch * scale), W // (patch * scale)
    g = x.view(B, Hq, Wq, C)
    h = gelu_erf(conv3x3_nhwc(g, sd[prefix + "seg_head.0.weight"], sd[prefix + "seg_head.0.bias"]))
    w2 = sd[prefix + "seg_head.2.weight"].view(1, -1)
    lo = (h @ w2.t()).squeeze(-1) + sd[prefix + "seg_head.2.bias"]     # [B,Hq,Wq]
    up = bilinear_up(lo, patch * scale)
    return up[:, None, :H, :W]


def upscaling_head(sd: SD, prefix: str, x: Tensor, padded_res, patch: int = 2):
    B, N, C = x.shape
    res = (padded_res[0] // patch, padded_res[1] // patch)
    for i in range(2):
        x, res = patch_expanding(sd, f"{prefix}ups.{i}.", x, res)
        x = basic_layer(sd, f"{prefix}swin_blocks.{i}.", x, res, 2, 3)
    g = x.view(B, res[0], res[1], C // 4)
    h = gelu_erf(conv3x3_nhwc(g, sd[prefix + "reconstruction.0.weight"], sd[prefix + "reconstruction.0.bias"]))
    w2 = sd[prefix + "reconstruction.2.weight"]
    out = h @ w2.view(w2.shape[0], -1).t() + sd[prefix + "reconstruction.2.bias"]
    return out.permute(0, 3, 1, 2).contiguous()


# ----------------------------------------------------------------------------
# cross attention  (SwinWNet.py:778-797; nn.MultiheadAttention batch_first, 3 heads)
# ----------------------------------------------------------------------------
def cross_attention_block(sd: SD, prefix: str, q: Tensor, kv: Tensor, num_heads: int = 3) -> Tensor:
    B, Lq, C = q.shape
    Lk = kv.shape[1]
    hd = C // num_heads
    qn = layer_norm(q, sd[prefix + "norm_q.weight"], sd[prefix + "norm_q.bias"])
    kn = layer_norm(kv, sd[prefix + "norm_kv.weight"], sd[prefix + "norm_kv.bias"])
    Wi, bi = sd[prefix + "attn.in_proj_weight"], sd[prefix + "attn.in_proj_bias"]
    Q = linear(qn, Wi[:C], bi[:C]).view(B, Lq, num_heads, hd).permute(0, 2, 1, 3)
    K = linear(kn, Wi[C:2 * C], bi[C:2 * C]).view(B, Lk, num_heads, hd).permute(0, 2, 1, 3)
    V = linear(kn, Wi[2 * C:], bi[2 * C:]).view(B, Lk, num_heads, hd).permute(0, 2, 1, 3)
    A = torch.softmax((Q @ K.transpose(-1, -2)) * (hd ** -0.5), dim=-1)
    o = (A @ V).permute(0, 2, 1, 3).reshape(B, Lq, C)
    o = linear(o, sd[prefix + "attn.out_proj.weight"], sd[prefix + "attn.out_proj.bias"])
    return q + sd[prefix + "gamma"] * o


def multi_scale_cross_attention(sd: SD, prefix: str, targets, sources):
    return [cross_attention_block(sd, f"{prefix}blocks.{i}.", t, s) for i, (t, s) in enumerate(zip(targets, sources))]


# ----------------------------------------------------------------------------
# model-level entry points  (SwinWNet.py:886-957, 574-592, 740-761)
# ----------------------------------------------------------------------------
DEPTHS = (2, 2, 2, 2)        # configuration of the shipped checkpoints
HEADS = (3, 6, 12, 24)


def segment_1(sd: SD, x: Tensor, depths=DEPTHS, heads=HEADS):
    t, pres = patch_embed(sd, "patch_embed.", x, 1)
    res = (pres[0] // 2, pres[1] // 2)
    skips, rl, bres = encoder(sd, "segmentator_encoder.", t, res, depths, heads)
    xb = bottleneck(sd, "segmentator_bottleneck.", skips[-1], bres, heads[-1])
    xd, _ = decoder(sd, "segmentator_decoder.", xb, bres, skips, rl, depths, heads)
    return segmentation_head(sd, "segmentator_head.", xd, pres, 1), skips


def upscale(sd: SD, x: Tensor, skips_seg: List[Tensor], depths=DEPTHS, heads=HEADS):
    rH, rW = x.shape[2] * 2, x.shape[3] * 2
    t, pres = patch_embed(sd, "patch_embed.", x, 1)
    res = (pres[0] // 2, pres[1] // 2)
    skips, rl, bres = encoder(sd, "upscaler_encoder.", t, res, depths, heads)
    ca = multi_scale_cross_attention(sd, "ca_seg_to_sr.", [skips[-2], skips[-1]], [skips_seg[-2], skips_seg[-1]])
    skips[-2], skips[-1] = ca
    xb = bottleneck(sd, "upscaler_bottleneck.", skips[-1], bres, heads[-1])
    xd, _ = decoder(sd, "upscaler_decoder.", xb, bres, skips, rl, depths, heads)
    up = upscaling_head(sd, "upscaler_head.", xd, pres)
    return up[:, :, :rH, :rW], skips


def segment_2(sd: SD, x: Tensor, skips_sr: List[Tensor], depths=DEPTHS, heads=HEADS):
    t, pres = patch_embed(sd, "patch_embed.", x, 2)
    res = (pres[0] // 4, pres[1] // 4)
    skips, rl, bres = encoder(sd, "segmentator_encoder.", t, res, depths, heads)
    ca = multi_scale_cross_attention(sd, "ca_sr_to_seg.", [skips[-2], skips[-1]], [skips_sr[-2], skips_sr[-1]])
    skips[-2], skips[-1] = ca
    xb = bottleneck(sd, "segmentator_bottleneck.", skips[-1], bres, heads[-1])
    xd, _ = decoder(sd, "segmentator_decoder.", xb, bres, skips, rl, depths, heads)
    return segmentation_head(sd, "segmentator_head.", xd, pres, 2), skips


def swin_unet(sd: SD, x: Tensor, depths=DEPTHS, heads=HEADS):
    t, pres = patch_embed(sd, "patch_embed.", x, 1)
    res = (pres[0] // 2, pres[1] // 2)
    skips, rl, bres = encoder(sd, "encoder.", t, res, depths, heads)
    xb = bottleneck(sd, "bottleneck.", skips[-1], bres, heads[-1])
    xd, _ = decoder(sd, "decoder.", xb, bres, skips, rl, depths, heads)
    return segmentation_head(sd, "head.", xd, pres, 1)


def swin_unet_sr(sd: SD, x: Tensor, depths=DEPTHS, heads=HEADS):
    rH, rW = x.shape[2] * 2, x.shape[3] * 2
    t, pres = patch_embed(sd, "patch_embed.", x, 1)
    res = (pres[0] // 2, pres[1] // 2)
    skips, rl, bres = encoder(sd, "encoder.", t, res, depths, heads)
    xb = bottleneck(sd, "bottleneck.", skips[-1], bres, heads[-1])
    xd, _ = decoder(sd, "decoder.", xb, bres, skips, rl, depths, heads)
    return upscaling_head(sd, "head.", xd, pres)[:, :, :rH, :rW]


# ----------------------------------------------------------------------------
# ST inference pipeline glue  (ST_Inference_Pipline.py:32-136)
# ----------------------------------------------------------------------------
def ensure_2ch(x: Tensor) -> Tensor:
    return x if x.size(1) == 2 else torch.cat([x, torch.sqrt(torch.abs(x))], dim=1)


def normalize_piecewise(x: Tensor, threshold: float = 0.01, eps: float = 1e-6):
    lo = x.amin(dim=(2, 3), keepdim=True)
    hi = x.amax(dim=(2, 3), keepdim=True)
    x01 = (x - lo) / (hi - lo + eps)
    return torch.where(x01 > threshold, torch.log1p(x01), x01), (lo, hi, threshold)


def denormalize_piecewise(y: Tensor, params, eps: float = 1e-6) -> Tensor:
    lo, hi, threshold = params
    x01 = torch.where(y > threshold, torch.expm1(y), y)
    return x01 * (hi - lo + eps) + lo


def st_pipeline(sd: SD, images: Tensor, depths=DEPTHS, heads=HEADS, two_channel: bool = True) -> Dict[str, Tensor]:
    """The 8 stages of SwinWNetInference.__call__ (ST_Inference_Pipline.py:73-136).

    two_channel=False is the manual call pattern used for the diffraction-only model
    (no ensure_2ch)."""
    out: Dict[str, Tensor] = {}
    if two_channel:
        images = ensure_2ch(images)
    out["images"] = images
    seg, skips_seg = segment_1(sd, images, depths, heads)
    out["seg_lr_logits"] = seg
    out["seg_map_lr"] = torch.sigmoid(seg)
    out["images_masked_lr"] = images * out["seg_map_lr"]
    norm, params = normalize_piecewise(out["images_masked_lr"])
    out["norm"] = norm
    up, skips_sr = upscale(sd, norm, skips_seg, depths, heads)
    out["upscaled_norm"] = up
    out["upscaled_denorm"] = denormalize_piecewise(up, params)
    seg_hr, _ = segment_2(sd, out["upscaled_denorm"], skips_sr, depths, heads)
    out["seg_hr_logits"] = seg_hr
    out["seg_map_hr"] = torch.sigmoid(seg_hr)
    out["images_masked_hr"] = out["upscaled_denorm"] * out["seg_map_hr"]
    return out


# ----------------------------------------------------------------------------
# helpers shared by tests / bench live in the neutral module benchdata.py (re-exported here)
# ----------------------------------------------------------------------------
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from benchdata import make_state_dict, surrogate_state_dict, synthetic_diffractions  # noqa: E402,F401
