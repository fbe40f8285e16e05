"""Differentiable fp32 restatements of the leaf operators of the hot path, in plain torch (ATen).

Used ONLY by the backward pass of ``autograd.py``: the forward of every operator runs in the sm_100a kernels; its
backward re-evaluates the operator here on the saved inputs with autograd enabled and back-propagates through ATen
("correct first": SURVEY.md §8 f-3 ranks hand-written backward kernels after the forward bar).
Nothing here is ever executed by an inference call, and there is no CPU path: the tensors are CUDA tensors.

Semantics follow the reference module by module (file:line = /root/reference/SwinWNet.py).
"""
import torch
import torch.nn.functional as F

WS = 5


def _rel_index(device):
    t = torch.arange(WS * WS, device=device)
    y, x = t // WS, t % WS
    return ((y[:, None] - y[None, :] + WS - 1) * (2 * WS - 1) + (x[:, None] - x[None, :] + WS - 1)).reshape(-1)


def patch_embed(x, w, b, nw, nb, scale=1, patch=2):
    """ScaleAwarePatchEmbed (:53-82): 2x2 conv, stride 2s, dilation s, flatten, LayerNorm(48)."""
    H, W = x.shape[-2:]
    pad_h = (patch * scale - H % patch * scale) % patch * scale      # the reference's literal precedence (:70-71)
    pad_w = (patch * scale - W % patch * scale) % patch * scale
    if pad_h or pad_w:
        x = F.pad(x, (0, pad_w, 0, pad_h))
    y = F.conv2d(x, w, b, stride=patch * scale, dilation=scale)
    return F.layer_norm(y.flatten(2).transpose(1, 2), (w.shape[0],), nw, nb)


def window_msa(xn, res, heads, qkv_w, qkv_b, table, proj_w, proj_b):
    """W-MSA on post-norm tokens, shift 0 (:183-209, window partition/reverse :86-121; zero pad AFTER the norm)."""
    B, L, C = xn.shape
    H, W = res
    hd = C // heads
    Hp, Wp = -(-H // WS) * WS, -(-W // WS) * WS
    g = F.pad(xn.view(B, H, W, C), (0, 0, 0, Wp - W, 0, Hp - H))
    ny, nx = Hp // WS, Wp // WS
    win = g.view(B, ny, WS, nx, WS, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, WS * WS, C)
    qkv = F.linear(win, qkv_w, qkv_b).view(-1, WS * WS, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = (qkv[0] * hd ** -0.5) @ qkv[1].transpose(-1, -2)
    att = att + table[_rel_index(table.device)].view(WS * WS, WS * WS, heads).permute(2, 0, 1)
    o = (att.softmax(-1) @ qkv[2]).transpose(1, 2).reshape(-1, WS * WS, C)
    o = F.linear(o, proj_w, proj_b).view(B, ny, nx, WS, WS, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    return o[:, :H, :W].reshape(B, L, C)


def swin_block(x, res, heads, n1w, n1b, qkv_w, qkv_b, table, proj_w, proj_b, n2w, n2b, w1, b1, w2, b2):
    """SwinTransformerBlock, shift 0 (:236-280): x + W-MSA(LN x), then x + MLP(LN x) with erf-GELU."""
    C = x.shape[-1]
    x = x + window_msa(F.layer_norm(x, (C,), n1w, n1b), res, heads, qkv_w, qkv_b, table, proj_w, proj_b)
    return x + F.linear(F.gelu(F.linear(F.layer_norm(x, (C,), n2w, n2b), w1, b1)), w2, b2)


def patch_merging(x, res, red_w, nw, nb):
    """PatchMerging (:289-316): zero-pad to even, concat [x00, x10, x01, x11], LayerNorm(4C), Linear(4C, 2C)."""
    B, L, C = x.shape
    H, W = res
    g = F.pad(x.view(B, H, W, C), (0, 0, 0, W % 2, 0, H % 2))
    cat = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], -1)
    return F.linear(F.layer_norm(cat.flatten(1, 2), (4 * C,), nw, nb), red_w)


def patch_expanding(x, res, target, exp_w, nw, nb):
    """PatchExpanding (:397-412) + crop_to_res (:414-424): Linear(C, 2C), pixel shuffle (2h+i, 2w+j), LayerNorm(C/2)."""
    B, L, C = x.shape
    H, W = res
    e = F.linear(x, exp_w).view(B, H, W, 2, 2, C // 2).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, C // 2)
    e = F.layer_norm(e, (C // 2,), nw, nb)
    Hs, Ws = target if target is not None else (2 * H, 2 * W)
    return e[:, :Hs, :Ws].reshape(B, Hs * Ws, C // 2)


def linear(x, w, b):
    return F.linear(x, w, b)


def segmentation_head(x, res, scale, w1, b1, w2, b2, patch=2):
    """SegmentationHead (:507-531): tokens -> NCHW, conv3x3 + GELU + conv1x1, bilinear x(2*scale), crop."""
    B, N, C = x.shape
    H, W = res
    up = patch * scale
    y = x.transpose(1, 2).reshape(B, C, H // up, W // up)
    y = F.conv2d(F.gelu(F.conv2d(y, w1, b1, padding=1)), w2, b2)
    y = F.interpolate(y, scale_factor=up, mode="bilinear", align_corners=False)
    return y[:, :, :H, :W]


def recon_tail(x, res, crop, w1, b1, w2, b2):
    """UpscalingHead tail (:682-688, crop :932): tokens -> NCHW, conv3x3 + GELU + conv1x1, crop."""
    B, N, C = x.shape
    H, W = res
    y = x.transpose(1, 2).reshape(B, C, H, W)
    y = F.conv2d(F.gelu(F.conv2d(y, w1, b1, padding=1)), w2, b2)
    return y[:, :, :crop[0], :crop[1]] if crop is not None else y


def cross_attention_block(q, kv, heads, in_w, in_b, out_w, out_b, nqw, nqb, nkw, nkb, gamma):
    """CrossAttentionBlock (:778-783): q + gamma * MHA(LN_q q, LN_kv kv, LN_kv kv)."""
    B, Lq, C = q.shape
    hd = C // heads
    qn, kn = F.layer_norm(q, (C,), nqw, nqb), F.layer_norm(kv, (C,), nkw, nkb)
    Q = F.linear(qn, in_w[:C], in_b[:C]).view(B, Lq, heads, hd).transpose(1, 2)
    K = F.linear(kn, in_w[C:2 * C], in_b[C:2 * C]).view(B, -1, heads, hd).transpose(1, 2)
    V = F.linear(kn, in_w[2 * C:], in_b[2 * C:]).view(B, -1, heads, hd).transpose(1, 2)
    o = ((Q @ K.transpose(-1, -2) * hd ** -0.5).softmax(-1) @ V).transpose(1, 2).reshape(B, Lq, C)
    return q + gamma * F.linear(o, out_w, out_b)
