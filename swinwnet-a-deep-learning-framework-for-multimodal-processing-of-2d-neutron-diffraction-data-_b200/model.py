"""Drop-in SwinWNet / SwinUNet / SwinUNetSR modules whose forward runs on the sm_100a kernels.

API contract: same class names, constructor signatures, sub-module attribute tree and
parameter names as /root/reference/SwinWNet.py (so ``state_dict`` / ``load_state_dict(strict=True)``
interoperate with the shipped ``models/*.pth``), same ``segment_1 / upscale / segment_2 / forward``
methods, same fp32 outputs and crops.  The sub-modules below are *parameter containers* (stock
``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Conv2d`` / ``nn.MultiheadAttention`` objects are used purely to
own identically named parameters); none of their ``forward`` methods is ever called.  The math is
lowered onto the C ABI in ``ops.py``:

    norm1+qkv            -> swn_rowgemm (LayerNorm prologue, bf16 epilogue)            SwinWNet.py:242,185
    W-MSA core           -> swn_window_attention (window/pad/roll/mask = index math)   SwinWNet.py:86-149,188-206
    proj + shortcut      -> swn_rowgemm (bf16 A, residual epilogue)                    SwinWNet.py:207,277
    norm2+MLP+shortcut   -> swn_mlp (fused fc1/GELU/fc2, hidden stays on chip)         SwinWNet.py:278
    PatchMerging         -> swn_rowgemm (2x2 gather + LayerNorm prologue)              SwinWNet.py:289-316
    PatchExpanding+crop  -> swn_rowgemm (pixel-shuffle + LayerNorm scatter epilogue)   SwinWNet.py:397-424
    skip concat          -> expand epilogue writes the left half, swn_copy_cols the right  SwinWNet.py:483
    CrossAttentionBlock  -> 2x swn_rowgemm (norm_q/norm_kv + in_proj) + swn_cross_attention
                            + swn_rowgemm (out_proj, q + gamma*o epilogue)             SwinWNet.py:778-783
    patch embed / heads  -> swn_patch_embed / swn_seg_head / swn_recon_head            SwinWNet.py:53-82,507-531,682-688

With autograd enabled (the reference trainers: Segmentator_pretrain.py:185, FullModel_supervised_trainer.py:231-288) every
leaf operator goes through ``autograd.KernelOp``: the forward is the same kernel lowering, the backward re-evaluates
the operator with its fp32 torch restatement (``torch_ref.py``) on the saved inputs.  Under ``torch.no_grad()`` nothing of
that is touched (in-place ping-pong buffers, no saved tensors).
"""
import os

import torch
import torch.nn as nn

from . import autograd, ops, packing, torch_ref

WINDOW = 5
FUSED_BLOCK = True   # route shift-0 blocks of the widths below through the single-kernel paths (csrc/swin_fused.cu)
FUSED_WHOLE = {(12, 3), (24, 3), (48, 3), (48, 6)}    # (C, heads) instances of swin_fused_kernel (whole block)
FUSED_ATTN = {(96, 3), (96, 6)}                       # instances of swin_attn_stream_kernel (attention half)
# (C, heads) that run one warp per window on mma.sync register fragments (csrc/swin_warp.cu) instead of swin_fused_kernel;
# SWN_WARP_BLOCK=0 keeps the tcgen05 block kernel for A/B measurements
WARP_BLOCK = {(12, 3), (24, 3), (48, 3), (48, 6)} if os.environ.get("SWN_WARP_BLOCK", "1") != "0" else set()
WARP_LAYER = os.environ.get("SWN_WARP_LAYER", "1") != "0"    # all blocks of such a BasicLayer in one launch


def _check_infer(x):
    if not x.is_cuda:
        raise RuntimeError("swinwnet_b200: forward needs CUDA tensors (B200); there is no CPU fallback")


def _grad():
    """autograd path?  (torch.autograd.Function.forward runs with grad mode off, so the kernel lowering below is what
    KernelOp.forward executes)"""
    return torch.is_grad_enabled()


class _PackCache:
    """Derived bf16 weight images, rebuilt when any source parameter changes (version / storage / device)."""

    def __init__(self):
        self._key, self._val = None, None

    def get(self, params, builder):
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in params if p is not None)
        if key != self._key:
            self._val, self._key = builder(), key
        return self._val


def _f32(t):
    return t.detach().float().contiguous()


# =============================================================================================
# building blocks (containers + kernel lowering)
# =============================================================================================
class ScaleAwarePatchEmbed(nn.Module):
    def __init__(self, patch_size=2, in_chans=1, embed_dim=48):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, dilation=1, padding=0,
                              bias=True)
        self.norm = nn.LayerNorm(embed_dim)

    def forward(self, x, scale_factor=1):
        _check_infer(x)
        if _grad():
            ps, s = self.patch_size, scale_factor
            H, W = x.shape[-2:]
            padded = (H + (ps * s - H % ps * s) % ps * s, W + (ps * s - W % ps * s) % ps * s)
            out = autograd.op(lambda x_, *p: self.forward(x_, scale_factor)[0],
                              lambda x_, *p: torch_ref.patch_embed(x_, *p, scale=scale_factor, patch=ps),
                              x, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias)
            return out, padded
        if self.patch_size != 2 or self.proj.out_channels != 48:
            raise RuntimeError("swinwnet_b200: kernels are built for patch_size=2, embed_dim=48")
        B, C, H, W = x.shape
        if C != self.proj.in_channels:
            raise RuntimeError(f"Given groups=1, weight of size {list(self.proj.weight.shape)}, expected input"
                               f"{list(x.shape)} to have {self.proj.in_channels} channels, but got {C} channels instead")
        ps, s = self.patch_size, scale_factor
        # the reference's pad formula, literal precedence (SwinWNet.py:70-71)
        pad_h = (ps * s - H % ps * s) % ps * s
        pad_w = (ps * s - W % ps * s) % ps * s
        Hn, Wn = H + pad_h, W + pad_w
        stride = ps * s
        Ho = (Hn - s * (ps - 1) - 1) // stride + 1
        Wo = (Wn - s * (ps - 1) - 1) // stride + 1
        out = torch.empty(B, Ho * Wo, 48, device=x.device, dtype=torch.float32)
        ops.patch_embed(x.float().contiguous(), _f32(self.proj.weight), _f32(self.proj.bias), _f32(self.norm.weight),
                        _f32(self.norm.bias), out, B, C, H, W, Ho, Wo, s)
        return out, (Hn, Wn)


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        t = torch.arange(window_size * window_size)
        y, x = t // window_size, t % window_size
        idx = (y[:, None] - y[None, :] + window_size - 1) * (2 * window_size - 1) + (x[:, None] - x[None, :] + window_size - 1)
        self.register_buffer("relative_position_index", idx.long())
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.normal_(self.relative_position_bias_table, std=.02)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, num_heads, window_size=5, shift_size=0, mlp_ratio=4., qkv_bias=True, drop=0., attn_drop=0.,
                 drop_path=0.):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.shift_size, self.mlp_ratio = dim, num_heads, window_size, shift_size, mlp_ratio
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, window_size=window_size, num_heads=num_heads, qkv_bias=qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        hidden = int(dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(drop), nn.Linear(hidden, dim), nn.Dropout(drop))
        self._cache = _PackCache()
        self._cache_fused = _PackCache()
        self._cache_warp = _PackCache()
        if drop or attn_drop or drop_path:
            raise NotImplementedError("swinwnet_b200: dropout/drop_path > 0 is a training feature (not implemented)")

    def _packed(self):
        a, C = self.attn, self.dim
        src = [a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, self.mlp[0].weight, self.mlp[0].bias,
               self.mlp[3].weight, self.mlp[3].bias, self.norm1.weight, self.norm1.bias, self.norm2.weight,
               self.norm2.bias, a.relative_position_bias_table]

        def build():
            if a.qkv.bias is None:
                raise NotImplementedError("swinwnet_b200: qkv_bias=False is not supported")
            if self.window_size != WINDOW or int(C * self.mlp_ratio) != 4 * C:
                raise NotImplementedError("swinwnet_b200: kernels are built for window_size=5, mlp_ratio=4")
            d = {}
            # N-chunk of the qkv GEMM: at K = 192 the weight tiles of a 192-column chunk (24 KB) leave room for only two ring
            # stages next to the resident A tile; 96-column chunks (5 stages) measured 20 % faster (0.63 -> 0.51 ms)
            nv = packing.choose_chunk(3 * C, 128 if C <= 192 else 256)
            d["qkv"] = packing.pack_rowgemm(a.qkv.weight, a.qkv.bias, nv) + (nv,)
            nv = packing.choose_chunk(C, 256)
            d["proj"] = packing.pack_rowgemm(a.proj.weight, a.proj.bias, nv) + (nv,)
            HC, TR = ops.mlp_config(C)
            d["mlp"] = packing.pack_mlp(self.mlp[0].weight, self.mlp[3].weight, self.mlp[3].bias, HC, TR)
            for k, p in (("qkv_b", a.qkv.bias), ("b1", self.mlp[0].bias), ("n1w", self.norm1.weight), ("n1b", self.norm1.bias),
                         ("n2w", self.norm2.weight), ("n2b", self.norm2.bias), ("tab", a.relative_position_bias_table)):
                d[k] = _f32(p)
            d["tabfrag"] = packing.rel_pos_bias_fragments(d["tab"], 1.4426950408889634).reshape(-1).contiguous()
            return d
        return self._cache.get(src, build)

    def _packed_fused(self):
        a = self.attn
        src = [self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.relative_position_bias_table,
               a.proj.weight, a.proj.bias, self.norm2.weight, self.norm2.bias, self.mlp[0].weight, self.mlp[0].bias,
               self.mlp[3].weight, self.mlp[3].bias]

        def build():
            if a.qkv.bias is None:
                raise NotImplementedError("swinwnet_b200: qkv_bias=False is not supported")
            if self.window_size != WINDOW or int(self.dim * self.mlp_ratio) != 4 * self.dim:
                raise NotImplementedError("swinwnet_b200: kernels are built for window_size=5, mlp_ratio=4")
            return packing.pack_fused_block(*src, self.num_heads)
        return self._cache_fused.get(src, build)

    def _packed_warp(self):
        a = self.attn
        src = [self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.relative_position_bias_table,
               a.proj.weight, a.proj.bias, self.norm2.weight, self.norm2.bias, self.mlp[0].weight, self.mlp[0].bias,
               self.mlp[3].weight, self.mlp[3].bias]

        def build():
            if a.qkv.bias is None:
                raise NotImplementedError("swinwnet_b200: qkv_bias=False is not supported")
            if self.window_size != WINDOW or int(self.dim * self.mlp_ratio) != 4 * self.dim:
                raise NotImplementedError("swinwnet_b200: kernels are built for window_size=5, mlp_ratio=4")
            return packing.pack_warp_block(*src, self.num_heads)
        return self._cache_warp.get(src, build)

    def _packed_attn_stream(self):
        a = self.attn
        src = [self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.relative_position_bias_table,
               a.proj.weight, a.proj.bias]
        return self._cache_fused.get(src, lambda: packing.pack_fused_attn_stream(*src, self.num_heads))

    def run(self, x, resolution, out, tmp=None):
        """x [B,L,C] fp32 -> out (may alias x).  `tmp` is a scratch tensor of the same shape: the attention half writes
        x + attn into it and the MLP half reads it, so no kernel ever runs in place (a thread that read-modify-writes its
        row in 16-byte pieces invalidates its own L1 lines: measured 2-6x slower than the out-of-place form)."""
        B, L, C = x.shape
        H, W = resolution
        assert L == H * W, "input feature has wrong size"
        if FUSED_BLOCK and (C, self.num_heads) in FUSED_WHOLE and self.shift_size == 0 and x.data_ptr() != out.data_ptr():
            if (C, self.num_heads) in WARP_BLOCK:
                # C = 12 / 24: one warp per window, everything in mma.sync register fragments (csrc/swin_warp.cu)
                Wpk, fpk = self._packed_warp()
                ops.swin_block_warp(x, out, B, H, W, C, self.num_heads, self.norm1.eps, Wpk, fpk)
                return out
            # narrow layers (C = 48, and 12 / 24 when opted out above): the whole block is one tcgen05 kernel (csrc/swin_fused.cu)
            Wpk, fpk = self._packed_fused()
            ops.swin_block_fused(x, out, B, H, W, C, self.num_heads, self.norm1.eps, Wpk, fpk, True)
            return out
        if C in (12, 24) and self.num_heads == 3 and self.window_size == WINDOW and int(C * self.mlp_ratio) == 4 * C \
                and self.attn.qkv.bias is not None:
            # narrow UpscalingHead layers: the whole block is one fp32 kernel (csrc/small_block.cu)
            a = self.attn
            params = [_f32(t) for t in (self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias,
                                        a.relative_position_bias_table, a.proj.weight, a.proj.bias, self.norm2.weight,
                                        self.norm2.bias, self.mlp[0].weight, self.mlp[0].bias, self.mlp[3].weight,
                                        self.mlp[3].bias)]
            ops.swin_block_small(x, out, B, H, W, C, self.num_heads, self.shift_size, self.norm1.eps, params)
            return out
        pk = self._packed()
        M = B * L
        if FUSED_BLOCK and (C, self.num_heads) in FUSED_ATTN and self.shift_size == 0:
            # attention half as one streamed-weight tcgen05 kernel (csrc/swin_fused.cu), MLP half as before
            if tmp is None:
                tmp = torch.empty_like(out)
            Wpk, fpk = self._packed_attn_stream()
            ops.swin_block_fused(x, tmp, B, H, W, C, self.num_heads, self.norm1.eps, Wpk, fpk, False)
            Wm, b2p = pk["mlp"]
            ops.mlp(tmp, out, M, C, pk["n2w"], pk["n2b"], Wm, pk["b1"], b2p, self.norm2.eps)
            return out
        qkv = torch.empty(M, 3 * C, device=x.device, dtype=ops.operand_dtype())
        Wp, bp, NT, nch, nv = pk["qkv"]
        ops.rowgemm(A=x, a_mode=ops.A_F32_LN, M=M, K=C, lda=C, ln_w=pk["n1w"], ln_b=pk["n1b"], ln_eps=self.norm1.eps,
                    Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_BF16, bias=bp, out=qkv, ldo=3 * C)
        att = torch.empty(M, C, device=x.device, dtype=ops.operand_dtype())
        ops.window_attention(qkv, att, pk["qkv_b"], pk["tab"], B, H, W, C, self.num_heads, self.shift_size, pk["tabfrag"])
        Wp, bp, NT, nch, nv = pk["proj"]
        if tmp is None:
            tmp = torch.empty_like(out)
        ops.rowgemm(A=att, a_mode=ops.A_BF16, M=M, K=C, lda=C, Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_F32,
                    bias=bp, out=tmp, ldo=C, res=x, ldres=C)
        Wm, b2p = pk["mlp"]
        ops.mlp(tmp, out, M, C, pk["n2w"], pk["n2b"], Wm, pk["b1"], b2p, self.norm2.eps)
        return out

    def _params(self):
        a = self.attn
        return (self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.relative_position_bias_table, a.proj.weight,
                a.proj.bias, self.norm2.weight, self.norm2.bias, self.mlp[0].weight, self.mlp[0].bias, self.mlp[3].weight,
                self.mlp[3].bias)

    def forward(self, x, resolution):
        _check_infer(x)
        if _grad():
            if self.shift_size:
                raise NotImplementedError("swinwnet_b200: shift_size > 0 has no backward (the reference cannot run it either, "
                                          "SwinWNet.py:147)")
            res, nh = tuple(resolution), self.num_heads
            return autograd.op(lambda x_, *p: self.forward(x_, res), lambda x_, *p: torch_ref.swin_block(x_, res, nh, *p),
                               x, *self._params())
        x = x.float().contiguous()
        return self.run(x, resolution, torch.empty_like(x))


class BasicLayer(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size=5, mlp_ratio=4., qkv_bias=True, drop=0., attn_drop=0., drop_path=0.):
        super().__init__()
        self.dim, self.depth = dim, depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim=dim, num_heads=num_heads, window_size=window_size, shift_size=0,  # SwinWNet.py:328
                                 mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop, attn_drop=attn_drop, drop_path=drop_path)
            for _ in range(depth)])
        self._cache_warp = _PackCache()

    def run(self, x, resolution, inplace=False):
        """returns the layer output; with inplace=True the caller gives up x (it may be overwritten or returned)."""
        if _grad():                       # autograd path: no buffer reuse, one KernelOp per block
            for blk in self.blocks:
                x = blk(x, resolution)
            return x
        if (FUSED_BLOCK and WARP_LAYER and 1 <= len(self.blocks) <= (4 if self.dim < 48 else 3) and x.shape[-1] == self.dim
                and all((b.dim, b.num_heads) in WARP_BLOCK and b.shift_size == 0 for b in self.blocks)):
            # C = 12 / 24: every block of the layer uses the same (unshifted) window partition -> ONE launch carries the rows
            # of a window through all blocks in registers (csrc/swin_warp.cu)
            B, L, C = x.shape
            H, W = resolution
            assert L == H * W, "input feature has wrong size"
            packs = [b._packed_warp() for b in self.blocks]
            Wpk, fpk = self._cache_warp.get([t for pk in packs for t in pk],
                                            lambda: (torch.cat([pk[0] for pk in packs]), torch.cat([pk[1] for pk in packs])))
            out = torch.empty_like(x)
            ops.swin_block_warp(x, out, B, H, W, C, self.blocks[0].num_heads, self.blocks[0].norm1.eps, Wpk, fpk, len(self.blocks))
            return out
        if FUSED_BLOCK and all((b.dim, b.num_heads) in FUSED_WHOLE and b.shift_size == 0 for b in self.blocks):
            # single-kernel blocks never run in place: ping-pong between two buffers
            spare = None
            for i, blk in enumerate(self.blocks):
                dst = spare if spare is not None else torch.empty_like(x)
                blk.run(x, resolution, dst)
                spare = x if (inplace or i > 0) else None
                x = dst
            return x
        out = x if inplace else torch.empty_like(x)
        tmp = torch.empty_like(x) if (len(self.blocks) and x.shape[-1] not in (12, 24)) else None
        for i, blk in enumerate(self.blocks):
            blk.run(x if i == 0 else out, resolution, out, tmp)
        return out if len(self.blocks) else x

    def forward(self, x, resolution):
        _check_infer(x)
        return self.run(x if _grad() else x.float().contiguous(), resolution)


class PatchMerging(nn.Module):
    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)
        self._cache = _PackCache()

    def forward(self, x, resolution):
        _check_infer(x)
        B, L, C = x.shape
        H, W = resolution
        assert L == H * W, "input feature has wrong size"
        Ho, Wo = (H + 1) // 2, (W + 1) // 2
        if _grad():
            res = (H, W)
            out = autograd.op(lambda x_, *p: self.forward(x_, res)[0], lambda x_, *p: torch_ref.patch_merging(x_, res, *p),
                              x, self.reduction.weight, self.norm.weight, self.norm.bias)
            return out, (Ho, Wo)

        def build():
            nv = packing.choose_chunk(2 * C, 64 if 4 * C > 384 else 256)   # K=768: the A tile alone is 192 KB
            return packing.pack_rowgemm(self.reduction.weight, None, nv) + (nv, _f32(self.norm.weight), _f32(self.norm.bias))
        Wp, _, NT, nch, nv, nw, nb = self._cache.get([self.reduction.weight, self.norm.weight, self.norm.bias], build)
        out = torch.empty(B, Ho * Wo, 2 * C, device=x.device, dtype=torch.float32)
        ops.rowgemm(A=x, a_mode=ops.A_MERGE_LN, M=B * Ho * Wo, K=4 * C, lda=4 * C, ln_w=nw, ln_b=nb, ln_eps=self.norm.eps,
                    merge=(H, W, C, Ho, Wo), Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_F32, out=out, ldo=2 * C)
        return out, (Ho, Wo)


class PatchExpanding(nn.Module):
    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.expand = nn.Linear(dim, 2 * dim, bias=False)
        self.norm = norm_layer(dim // 2)
        self._cache = _PackCache()

    def run(self, x, resolution, target_res=None, out=None, ldo=None):
        """expand + pixel shuffle + LayerNorm, cropped to target_res, written to channels [0, C/2) of `out`
        whose row stride is `ldo` floats (lets the decoder write straight into its concat buffer)."""
        B, L, C = x.shape
        H, W = resolution
        assert L == H * W, "input feature has wrong size"
        Hs, Ws = target_res if target_res is not None else (2 * H, 2 * W)
        assert 2 * H >= Hs and 2 * W >= Ws
        Cg = C // 2
        if _grad():
            assert out is None, "the autograd path allocates its own outputs"
            res, tgt = (H, W), (Hs, Ws)
            y = autograd.op(lambda x_, *p: self.run(x_.float().contiguous(), res, tgt)[0],
                            lambda x_, *p: torch_ref.patch_expanding(x_, res, tgt, *p),
                            x, self.expand.weight, self.norm.weight, self.norm.bias)
            return y, (Hs, Ws)

        def build():
            return packing.pack_rowgemm(self.expand.weight, None, Cg) + (_f32(self.norm.weight), _f32(self.norm.bias))
        Wp, _, NT, nch, nw, nb = self._cache.get([self.expand.weight, self.norm.weight, self.norm.bias], build)
        if out is None:
            ldo = Cg
            out = torch.empty(B, Hs * Ws, Cg, device=x.device, dtype=torch.float32)
        ops.rowgemm(A=x, a_mode=ops.A_F32, M=B * L, K=C, lda=C, Wp=Wp, NT=NT, nchunks=nch, n_valid=Cg, e_mode=ops.E_EXPAND,
                    out=out, ldo=ldo, expand=(H, W, Hs, Ws), ln2_w=nw, ln2_b=nb, ln_eps=self.norm.eps)
        return out, (Hs, Ws)

    def forward(self, x, resolution):
        _check_infer(x)
        return self.run(x if _grad() else x.float().contiguous(), resolution)


class SwinEncoder(nn.Module):
    def __init__(self, embed_dim=48, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=5, mlp_ratio=4.,
                 qkv_bias=True, drop=0., attn_drop=0., drop_path=0.):
        super().__init__()
        self.layers, self.downs = nn.ModuleList(), nn.ModuleList()
        dim = embed_dim
        for i in range(len(depths) - 1):
            self.layers.append(BasicLayer(dim, depths[i], num_heads[i], window_size, mlp_ratio, qkv_bias, drop, attn_drop, drop_path))
            self.downs.append(PatchMerging(dim))
            dim *= 2
        self.layers.append(BasicLayer(dim, depths[-1], num_heads[-1], window_size, mlp_ratio, qkv_bias, drop, attn_drop, drop_path))

    def forward(self, x, resolution):
        _check_infer(x)
        skips, res_skips = [], []
        for i in range(len(self.layers) - 1):
            x = self.layers[i].run(x, resolution, inplace=i > 0)   # i>0: x is the fresh PatchMerging output
            skips.append(x)
            res_skips.append(resolution)
            x, resolution = self.downs[i](x, resolution)
        x = self.layers[-1].run(x, resolution, inplace=len(self.layers) > 1)
        skips.append(x)
        res_skips.append(resolution)
        return skips, res_skips, resolution


class Bottleneck(nn.Module):
    def __init__(self, dim, num_heads, window_size=5, mlp_ratio=4., qkv_bias=True, drop=0., attn_drop=0., drop_path=0.):
        super().__init__()
        self.layer = BasicLayer(dim, 2, num_heads, window_size, mlp_ratio, qkv_bias, drop, attn_drop, drop_path)

    def forward(self, x, resolution):
        _check_infer(x)
        return self.layer.run(x, resolution)


class SwinDecoder(nn.Module):
    def __init__(self, embed_dim=48, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=5, mlp_ratio=4.,
                 qkv_bias=True, drop=0., attn_drop=0., drop_path=0.):
        super().__init__()
        self.ups, self.swin_blocks, self.linears = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        dim = embed_dim * 8
        self.depths = depths[-2::-1]
        self.num_heads = num_heads[-2::-1]
        for i in range(len(depths) - 1):
            self.ups.append(PatchExpanding(dim=dim))
            self.swin_blocks.append(BasicLayer(dim, self.depths[i], self.num_heads[i], window_size, mlp_ratio, qkv_bias,
                                               drop, attn_drop, drop_path))
            self.linears.append(nn.Linear(dim, dim // 2))
            dim //= 2
        self._caches = [_PackCache() for _ in range(len(depths) - 1)]

    def forward(self, x, resolution, skips, skip_res_list):
        _check_infer(x)
        skips, skip_res_list = skips[-2::-1], skip_res_list[-2::-1]
        if _grad():
            for i in range(len(self.swin_blocks)):
                Hs, Ws = skip_res_list[i]
                xe, _ = self.ups[i].run(x, resolution, (Hs, Ws))                 # expand + crop_to_res
                cat = self.swin_blocks[i].run(torch.cat([xe, skips[i].float()], dim=-1), (Hs, Ws))
                lin = self.linears[i]
                x = autograd.op(lambda c_, *p, i=i: self._linear_infer(i, c_), lambda c_, w, b: torch_ref.linear(c_, w, b),
                                cat, lin.weight, lin.bias)
                resolution = (Hs, Ws)
            return x, resolution
        for i in range(len(self.swin_blocks)):
            B, L, C = x.shape                       # C = concat width of this stage
            Hs, Ws = skip_res_list[i]
            half = C // 2
            cat = torch.empty(B, Hs * Ws, C, device=x.device, dtype=torch.float32)
            self.ups[i].run(x, resolution, (Hs, Ws), out=cat, ldo=C)        # expanded -> channels [0, C/2)
            skip = skips[i].float().contiguous()
            assert skip.shape == (B, Hs * Ws, half)
            ops.copy_cols(skip, half, cat, half, C, B * Hs * Ws, half)       # skip -> channels [C/2, C)
            cat = self.swin_blocks[i].run(cat, (Hs, Ws), inplace=True)
            x = self._linear_infer(i, cat)
            resolution = (Hs, Ws)
        return x, resolution

    def _linear_infer(self, i, cat):
        """decoder stage linear 2c -> c on the raw fp32 stream (SwinWNet.py:489)"""
        lin = self.linears[i]
        cat = cat.float().contiguous()
        B, L, C = cat.shape
        half = C // 2

        def build():
            nv = packing.choose_chunk(half, 256)
            return packing.pack_rowgemm(lin.weight, lin.bias, nv) + (nv,)
        Wp, bp, NT, nch, nv = self._caches[i].get([lin.weight, lin.bias], build)
        x = torch.empty(B, L, half, device=cat.device, dtype=torch.float32)
        ops.rowgemm(A=cat, a_mode=ops.A_F32, M=B * L, K=C, lda=C, Wp=Wp, NT=NT, nchunks=nch, n_valid=nv,
                    e_mode=ops.E_F32, bias=bp, out=x, ldo=half)
        return x


class SegmentationHead(nn.Module):
    def __init__(self, embed_dim=48, patch_size=2):
        super().__init__()
        self.patch_size = patch_size
        self.seg_head = nn.Sequential(nn.Conv2d(embed_dim, embed_dim // 2, kernel_size=3, padding=1), nn.GELU(),
                                      nn.Conv2d(embed_dim // 2, 1, kernel_size=1))

    def forward(self, x, resolution, scale_factor=1):
        _check_infer(x)
        B, N, C = x.shape
        H, W = resolution
        up = self.patch_size * scale_factor
        Hq, Wq = H // up, W // up
        assert N == Hq * Wq and C == 48
        if _grad():
            res, c1, c2 = (H, W), self.seg_head[0], self.seg_head[2]
            return autograd.op(lambda x_, *p: self.forward(x_.float().contiguous(), res, scale_factor),
                               lambda x_, *p: torch_ref.segmentation_head(x_, res, scale_factor, *p, patch=self.patch_size),
                               x, c1.weight, c1.bias, c2.weight, c2.bias)
        lowres = torch.empty(B, Hq, Wq, device=x.device, dtype=torch.float32)
        Hout, Wout = min(H, Hq * up), min(W, Wq * up)
        out = torch.empty(B, 1, Hout, Wout, device=x.device, dtype=torch.float32)
        c1, c2 = self.seg_head[0], self.seg_head[2]
        ops.seg_head(x, _f32(c1.weight), _f32(c1.bias), _f32(c2.weight), _f32(c2.bias), lowres, out, B, Hq, Wq, up, Hout, Wout)
        return out


class UpscalingHead(nn.Module):
    def __init__(self, error_matrix=False, embed_dim=48, patch_size=2, window_size=5, num_heads=3, depth=2, mlp_ratio=4.,
                 qkv_bias=True, drop=0., attn_drop=0., drop_path=0.):
        super().__init__()
        self.patch_size = patch_size
        self.ups, self.swin_blocks = nn.ModuleList(), nn.ModuleList()
        for _ in range(2):
            self.ups.append(PatchExpanding(dim=embed_dim))
            self.swin_blocks.append(BasicLayer(embed_dim // 2, depth, num_heads, window_size, mlp_ratio, qkv_bias, drop,
                                               attn_drop, drop_path))
            embed_dim //= 2
        self.reconstruction = nn.Sequential(nn.Conv2d(embed_dim, embed_dim, kernel_size=3, padding=1), nn.GELU(),
                                            nn.Conv2d(embed_dim, 2 if error_matrix else 1, kernel_size=1))

    def forward(self, x, resolution, crop=None):
        _check_infer(x)
        B, N, C = x.shape
        res = (resolution[0] // self.patch_size, resolution[1] // self.patch_size)
        for i in range(2):
            x, res = self.ups[i].run(x, res)
            x = self.swin_blocks[i].run(x, res, inplace=True)
        return self._tail(x, res, crop)

    def _tail(self, x, res, crop):
        B = x.shape[0]
        Hh, Wh = res
        c1, c2 = self.reconstruction[0], self.reconstruction[2]
        Cout = c2.out_channels
        Hout, Wout = (min(crop[0], Hh), min(crop[1], Wh)) if crop is not None else (Hh, Wh)
        if _grad():
            r, cr = tuple(res), (Hout, Wout)
            return autograd.op(lambda x_, *p: self._tail(x_.float().contiguous(), r, cr),
                               lambda x_, *p: torch_ref.recon_tail(x_, r, cr, *p), x, c1.weight, c1.bias, c2.weight, c2.bias)
        out = torch.empty(B, Cout, Hout, Wout, device=x.device, dtype=torch.float32)
        ops.recon_head(x, _f32(c1.weight), _f32(c1.bias), _f32(c2.weight), _f32(c2.bias), out, B, Hh, Wh, Cout, Hout, Wout)
        return out


class CrossAttentionBlock(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.norm_q = nn.LayerNorm(dim)
        self.norm_kv = nn.LayerNorm(dim)
        self.attn = nn.MultiheadAttention(embed_dim=dim, num_heads=num_heads, batch_first=True)
        self.gamma = nn.Parameter(torch.zeros(1))
        self._cache = _PackCache()

    def forward(self, q, kv):
        _check_infer(q)
        B, Lq, C = q.shape
        Lk = kv.shape[1]
        a = self.attn
        if _grad():
            nh = a.num_heads
            return autograd.op(lambda q_, kv_, *p: self.forward(q_, kv_),
                               lambda q_, kv_, *p: torch_ref.cross_attention_block(q_, kv_, nh, *p),
                               q, kv, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias, self.norm_q.weight,
                               self.norm_q.bias, self.norm_kv.weight, self.norm_kv.bias, self.gamma)

        def build():
            Wi, bi = a.in_proj_weight, a.in_proj_bias
            nq, nkv, no = packing.choose_chunk(C, 256), packing.choose_chunk(2 * C, 256), packing.choose_chunk(C, 256)
            return dict(q=packing.pack_rowgemm(Wi[:C], bi[:C], nq) + (nq,), kv=packing.pack_rowgemm(Wi[C:], bi[C:], nkv) + (nkv,),
                        o=packing.pack_rowgemm(a.out_proj.weight, a.out_proj.bias, no) + (no,),
                        nqw=_f32(self.norm_q.weight), nqb=_f32(self.norm_q.bias), nkw=_f32(self.norm_kv.weight),
                        nkb=_f32(self.norm_kv.bias), gamma=_f32(self.gamma))
        pk = self._cache.get([a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias, self.norm_q.weight,
                              self.norm_q.bias, self.norm_kv.weight, self.norm_kv.bias, self.gamma], build)
        q = q.float().contiguous()
        kv = kv.float().contiguous()
        dev = q.device
        Qp = torch.empty(B * Lq, C, device=dev, dtype=ops.operand_dtype())
        KVp = torch.empty(B * Lk, 2 * C, device=dev, dtype=ops.operand_dtype())
        Wp, bp, NT, nch, nv = pk["q"]
        ops.rowgemm(A=q, a_mode=ops.A_F32_LN, M=B * Lq, K=C, lda=C, ln_w=pk["nqw"], ln_b=pk["nqb"], ln_eps=self.norm_q.eps,
                    Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_BF16, bias=bp, out=Qp, ldo=C)
        Wp, bp, NT, nch, nv = pk["kv"]
        ops.rowgemm(A=kv, a_mode=ops.A_F32_LN, M=B * Lk, K=C, lda=C, ln_w=pk["nkw"], ln_b=pk["nkb"], ln_eps=self.norm_kv.eps,
                    Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_BF16, bias=bp, out=KVp, ldo=2 * C)
        O = torch.empty(B * Lq, C, device=dev, dtype=ops.operand_dtype())
        ops.cross_attention(Qp, KVp, O, B, Lq, Lk, C, a.num_heads)
        out = torch.empty_like(q)
        Wp, bp, NT, nch, nv = pk["o"]
        ops.rowgemm(A=O, a_mode=ops.A_BF16, M=B * Lq, K=C, lda=C, Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_F32,
                    bias=bp, out=out, ldo=C, res=q, ldres=C, alpha=pk["gamma"])
        return out


class MultiScaleCrossAttention(nn.Module):
    def __init__(self, dims, heads):
        super().__init__()
        self.blocks = nn.ModuleList([CrossAttentionBlock(d, h) for d, h in zip(dims, heads)])

    def forward(self, target_skips, source_skips):
        return [blk(t, s) for blk, t, s in zip(self.blocks, target_skips, source_skips)]


# =============================================================================================
# models
# =============================================================================================
def _trunk(patch_embed, encoder, bottleneck, decoder, x, scale=1, cross=None):
    x_patch, padded_res = patch_embed(x, scale_factor=scale)
    ps = patch_embed.patch_size * scale
    resolution = (padded_res[0] // ps, padded_res[1] // ps)
    skips, skip_res, bott_res = encoder(x_patch, resolution)
    if cross is not None:
        ca, src = cross
        skips[-2], skips[-1] = ca([skips[-2], skips[-1]], [src[-2], src[-1]])
    x_b = bottleneck(skips[-1], bott_res)
    x_dec, _ = decoder(x_b, bott_res, skips, skip_res)
    return x_dec, padded_res, skips


class SwinUNet(nn.Module):
    def __init__(self, patch_size=2, in_chans=1, embed_dim=48, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                 window_size=5, mlp_ratio=4., qkv_bias=True, drop=0., attn_drop=0., drop_path=0.):
        super().__init__()
        self.patch_embed = ScaleAwarePatchEmbed(patch_size, in_chans, embed_dim)
        self.encoder = SwinEncoder(embed_dim, depths, num_heads, window_size, mlp_ratio, True, drop, attn_drop, drop_path)
        self.bottleneck = Bottleneck(embed_dim * 8, num_heads[-1], window_size)
        self.decoder = SwinDecoder(embed_dim, depths, num_heads, window_size, mlp_ratio, qkv_bias, drop, attn_drop, drop_path)
        self.head = SegmentationHead(embed_dim, patch_size)

    def forward(self, x):
        x_dec, padded_res, _ = _trunk(self.patch_embed, self.encoder, self.bottleneck, self.decoder, x)
        return self.head(x_dec, padded_res)


class SwinUNetSR(nn.Module):
    def __init__(self, patch_size=2, in_chans=1, embed_dim=48, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                 window_size=5, mlp_ratio=4., qkv_bias=True, drop=0., attn_drop=0., drop_path=0.):
        super().__init__()
        self.patch_embed = ScaleAwarePatchEmbed(patch_size, in_chans, embed_dim)
        self.encoder = SwinEncoder(embed_dim, depths, num_heads, window_size, mlp_ratio, True, drop, attn_drop, drop_path)
        self.bottleneck = Bottleneck(embed_dim * 8, num_heads[-1], window_size)
        self.decoder = SwinDecoder(embed_dim, depths, num_heads, window_size, mlp_ratio, qkv_bias, drop, attn_drop, drop_path)
        self.head = UpscalingHead(embed_dim=embed_dim, patch_size=patch_size, window_size=window_size, num_heads=3, depth=2,
                                  mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop, attn_drop=attn_drop, drop_path=drop_path)

    def forward(self, x):
        x_dec, padded_res, _ = _trunk(self.patch_embed, self.encoder, self.bottleneck, self.decoder, x)
        return self.head(x_dec, padded_res, crop=(x.size(2) * 2, x.size(3) * 2))


class SwinWNet(nn.Module):
    def __init__(self, patch_size=2, in_chans=1, error_matrix=False, embed_dim=48, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=5, mlp_ratio=4., qkv_bias=True, drop=0., attn_drop=0., drop_path=0.):
        super().__init__()
        self.patch_embed = ScaleAwarePatchEmbed(patch_size, in_chans + 1 if error_matrix else in_chans, embed_dim)
        enc = lambda: SwinEncoder(embed_dim, depths, num_heads, window_size, mlp_ratio, True, drop, attn_drop, drop_path)
        dec = lambda: SwinDecoder(embed_dim, depths, num_heads, window_size, mlp_ratio, qkv_bias, drop, attn_drop, drop_path)
        self.segmentator_encoder = enc()
        self.segmentator_bottleneck = Bottleneck(embed_dim * 8, num_heads[-1], window_size)
        self.segmentator_decoder = dec()
        self.segmentator_head = SegmentationHead(embed_dim, patch_size)
        self.ca_seg_to_sr = MultiScaleCrossAttention(dims=[embed_dim * 4, embed_dim * 8], heads=[3, 3])
        self.ca_sr_to_seg = MultiScaleCrossAttention(dims=[embed_dim * 4, embed_dim * 8], heads=[3, 3])
        self.upscaler_encoder = enc()
        self.upscaler_bottleneck = Bottleneck(embed_dim * 8, num_heads[-1], window_size)
        self.upscaler_decoder = dec()
        self.upscaler_head = UpscalingHead(error_matrix=error_matrix, embed_dim=embed_dim, patch_size=patch_size,
                                           window_size=window_size, num_heads=3, depth=2, mlp_ratio=mlp_ratio,
                                           qkv_bias=qkv_bias, drop=drop, attn_drop=attn_drop, drop_path=drop_path)

    def segment_1(self, x):
        x_dec, padded_res, skips = _trunk(self.patch_embed, self.segmentator_encoder, self.segmentator_bottleneck,
                                          self.segmentator_decoder, x)
        return self.segmentator_head(x_dec, padded_res), skips

    def upscale(self, x, skips_segmentator):
        x_dec, padded_res, skips = _trunk(self.patch_embed, self.upscaler_encoder, self.upscaler_bottleneck,
                                          self.upscaler_decoder, x, cross=(self.ca_seg_to_sr, skips_segmentator))
        return self.upscaler_head(x_dec, padded_res, crop=(x.size(2) * 2, x.size(3) * 2)), skips

    def segment_2(self, x, skips_upscaler):
        x_dec, padded_res, skips = _trunk(self.patch_embed, self.segmentator_encoder, self.segmentator_bottleneck,
                                          self.segmentator_decoder, x, scale=2, cross=(self.ca_sr_to_seg, skips_upscaler))
        return self.segmentator_head(x_dec, padded_res, scale_factor=2), skips
