"""Thin torch<->C-ABI glue: unwraps data_ptr()/current stream and calls libswinwnet_b200.so.
PyTorch is used for device memory and streams only; all math runs in the hand-written kernels."""
import ctypes

import torch

from . import _lib
from ._lib import RowGemmArgs

A_F32_LN, A_F32, A_BF16, A_MERGE_LN = 0, 1, 2, 3
E_BF16, E_F32, E_EXPAND = 0, 1, 2

def operand_dtype():
    """16-bit tensor-core operand dtype of the loaded library (fp16 by default, bf16 with -DSWN_OPERAND_BF16=1)."""
    return torch.bfloat16 if _lib.load().swn_operand_is_bf16() else torch.float16


LAUNCH_COUNT = 0  # kernels launched through this module (bench.py reports it as gpu_launches)
_LAUNCHES_PER_CALL = {"swin_block_small": 1, "swin_block_fused": 1, "swin_block_warp": 1, "rowgemm": 1,"mlp": 1, "window_attention": 1, "cross_attention": 1, "patch_embed": 1,
                      "seg_head": 2, "recon_head": 1, "copy_cols": 1, "sigmoid_mask": 1, "sigmoid_mask_mm": 2,
                      "normalize": 1, "dspace_histogram": 2, "adamw_multi": 1, "ensure_2ch": 1, "grad_bucket_copy": 1}


def _count(kind):
    global LAUNCH_COUNT
    LAUNCH_COUNT += _LAUNCHES_PER_CALL[kind]


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class _Launch:
    """Device guard of one C-ABI call.  The library launches on whatever device is current in the calling thread and the
    stream it is handed, so every wrapper (1) checks that all tensor arguments live on ONE CUDA device, (2) makes that
    device current for the duration of the call when it is not already (``SwinWNetInference(model, 'cuda:1')`` must work
    without a prior ``torch.cuda.set_device(1)``, like the reference), and (3) passes that device's current stream."""
    __slots__ = ("dev", "prev")

    def __init__(self, *ts):
        dev = None
        for t in ts:
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError("swinwnet_b200: tensors must live on a CUDA device (no CPU fallback exists)")
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise RuntimeError(f"swinwnet_b200: tensor arguments on different devices ({dev} vs {t.device})")
        self.dev, self.prev = dev, None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.dev.index is not None and self.dev.index != cur:
            self.prev = cur
            torch.cuda.set_device(self.dev)
        return ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def mlp_config(C):
    hc, tr = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.load().swn_mlp_config(C, ctypes.byref(hc), ctypes.byref(tr)), "swn_mlp_config")
    return hc.value, tr.value


def rowgemm(*, A, a_mode, M, K, lda, Wp, NT, nchunks, n_valid, e_mode, out, ldo, bias=None, ln_w=None, ln_b=None,
            ln_eps=1e-5, merge=None, res=None, ldres=0, alpha=None, expand=None, ln2_w=None, ln2_b=None):
    a = RowGemmArgs()
    a.A, a.a_mode, a.M, a.K, a.lda = _ptr(A), a_mode, M, K, lda
    a.ln_w, a.ln_b, a.ln_eps = _ptr(ln_w), _ptr(ln_b), ln_eps
    if merge is not None:
        a.gH, a.gW, a.gC, a.gHo, a.gWo = merge
    a.Wp, a.NT, a.nchunks, a.n_valid = _ptr(Wp), NT, nchunks, n_valid
    a.e_mode, a.bias, a.out, a.ldo = e_mode, _ptr(bias), _ptr(out), ldo
    a.res, a.ldres, a.alpha = _ptr(res), ldres, _ptr(alpha)
    if expand is not None:
        a.xH, a.xW, a.xHs, a.xWs = expand
    a.ln2_w, a.ln2_b = _ptr(ln2_w), _ptr(ln2_b)
    with _Launch(A, Wp, out, bias, ln_w, ln_b, res, alpha, ln2_w, ln2_b) as st:
        _lib.check(_lib.load().swn_rowgemm(ctypes.byref(a), st), "swn_rowgemm")
    _count("rowgemm")


def mlp(x, out, M, C, ln_w, ln_b, Wp, b1, b2p, eps=1e-5):
    with _Launch(x, out, Wp, ln_w, ln_b, b1, b2p) as st:
        _lib.check(_lib.load().swn_mlp(_ptr(x), _ptr(out), M, C, _ptr(ln_w), _ptr(ln_b), eps, _ptr(Wp), _ptr(b1), _ptr(b2p), st),
                   "swn_mlp")
    _count("mlp")


def swin_block_small(x, out, B, H, W, C, nH, shift, eps, params):
    """whole Swin block for C in {12, 24}; params = 13 fp32 device tensors (see include/swinwnet_b200.h)."""
    arr = (ctypes.c_void_p * 13)(*[p.data_ptr() for p in params])
    with _Launch(x, out, *params) as st:
        _lib.check(_lib.load().swn_swin_block_small(_ptr(x), _ptr(out), B, H, W, C, nH, shift, eps, arr, st),
                   "swn_swin_block_small")
    _count("swin_block_small")


def swin_block_fused(x, out, B, H, W, C, nH, eps, Wpk, fpk, do_mlp=True):
    """whole shift-0 Swin block (or its attention half) in one tcgen05 kernel, C <= 64; out must not alias x."""
    with _Launch(x, out, Wpk, fpk) as st:
        _lib.check(_lib.load().swn_swin_block_fused(_ptr(x), _ptr(out), B, H, W, C, nH, eps, _ptr(Wpk), _ptr(fpk),
                                                    int(do_mlp), st), "swn_swin_block_fused")
    _count("swin_block_fused")


def swin_block_warp(x, out, B, H, W, C, nH, eps, Wpk, fpk, depth=1):
    """`depth` consecutive shift-0 Swin blocks for C = 12 / 24 (3 heads): one warp per window on mma.sync register
    fragments (csrc/swin_warp.cu); Wpk / fpk = the blocks' packs concatenated; out must not alias x."""
    with _Launch(x, out, Wpk, fpk) as st:
        _lib.check(_lib.load().swn_swin_block_warp(_ptr(x), _ptr(out), B, H, W, C, nH, eps, _ptr(Wpk), _ptr(fpk), int(depth),
                                                   st), "swn_swin_block_warp")
    _count("swin_block_warp")


def set_phase_profile(buf):
    """profiling aid: int64 device tensor [grid, 16] (zeroed) that the fused block kernels add per-phase cycles to;
    None disables it."""
    _lib.check(_lib.load().swn_set_phase_profile(_ptr(buf)), "swn_set_phase_profile")


def window_attention(qkv, out, qkv_bias, table, B, H, W, C, nH, shift=0, bias_frags=None):
    """bias_frags (shift 0 only): packing.rel_pos_bias_fragments(table, log2 e), cached by the caller — saves the per-CTA
    rebuild of the bias images from the table"""
    with _Launch(qkv, out, qkv_bias, table, bias_frags) as st:
        if bias_frags is not None and shift == 0:
            _lib.check(_lib.load().swn_window_attention_frags(_ptr(qkv), _ptr(out), _ptr(qkv_bias), _ptr(table), _ptr(bias_frags),
                                                              B, H, W, C, nH, st), "swn_window_attention_frags")
        else:
            _lib.check(_lib.load().swn_window_attention(_ptr(qkv), _ptr(out), _ptr(qkv_bias), _ptr(table), B, H, W, C, nH, shift,
                                                        st), "swn_window_attention")
    _count("window_attention")


def cross_attention(q, kv, out, B, Lq, Lk, C, nH):
    with _Launch(q, kv, out) as st:
        _lib.check(_lib.load().swn_cross_attention(_ptr(q), _ptr(kv), _ptr(out), B, Lq, Lk, C, nH, st), "swn_cross_attention")
    _count("cross_attention")


def patch_embed(x, w, b, ln_w, ln_b, out, B, Cin, H, W, Ho, Wo, scale):
    with _Launch(x, out, w, b, ln_w, ln_b) as st:
        _lib.check(_lib.load().swn_patch_embed(_ptr(x), _ptr(w), _ptr(b), _ptr(ln_w), _ptr(ln_b), _ptr(out), B, Cin, H, W, Ho,
                                               Wo, scale, st), "swn_patch_embed")
    _count("patch_embed")


def seg_head(tok, w1, b1, w2, b2, lowres, out, B, Hq, Wq, up, Hout, Wout):
    with _Launch(tok, out, w1, b1, w2, b2, lowres) as st:
        _lib.check(_lib.load().swn_seg_head(_ptr(tok), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(lowres), _ptr(out), B, Hq,
                                            Wq, up, Hout, Wout, st), "swn_seg_head")
    _count("seg_head")


def recon_head(tok, w1, b1, w2, b2, out, B, Hh, Wh, Cout, Hout, Wout):
    with _Launch(tok, out, w1, b1, w2, b2) as st:
        _lib.check(_lib.load().swn_recon_head(_ptr(tok), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(out), B, Hh, Wh, Cout,
                                              Hout, Wout, st), "swn_recon_head")
    _count("recon_head")


def copy_cols(src, lds, dst_ptr_tensor, dst_col_offset, ldd, rows, cols):
    """dst[:, dst_col_offset : dst_col_offset+cols] = src[:, :cols] (row strides lds / ldd, in floats)."""
    dst = ctypes.c_void_p(dst_ptr_tensor.data_ptr() + 4 * dst_col_offset)
    with _Launch(src, dst_ptr_tensor) as st:
        _lib.check(_lib.load().swn_copy_cols(_ptr(src), lds, dst, ldd, rows, cols, st), "swn_copy_cols")
    _count("copy_cols")


def sigmoid_mask(img, seg, *, ensure_2ch, want_minmax):
    """returns (images, seg_map, masked, minmax|None) for the ST pipeline stages 1-3 / 7-8."""
    B, Cimg, H, W = img.shape
    # the reference multiplies images * sigmoid(seg) with broadcasting: a seg map of another size (e.g. the padded
    # [B,1,H+1,W+1] that segment_1 returns for odd H / W) is an error there (ST_Inference_Pipline.py:96), and here
    if tuple(seg.shape) != (B, 1, H, W):
        raise RuntimeError(f"The size of tensor a ({W}) must match the size of tensor b ({seg.shape[-1]}): images "
                           f"{tuple(img.shape)} vs segmentation map {tuple(seg.shape)}")
    if img.dtype != torch.float32 or seg.dtype != torch.float32 or not img.is_contiguous() or not seg.is_contiguous():
        raise RuntimeError("swinwnet_b200: sigmoid_mask needs contiguous fp32 tensors")
    Cout = 2 if (ensure_2ch and Cimg != 2) else Cimg
    images2 = torch.empty(B, Cout, H, W, device=img.device, dtype=torch.float32) if Cout != Cimg else None
    seg_map = torch.empty(B, 1, H, W, device=img.device, dtype=torch.float32)
    masked = torch.empty(B, Cout, H, W, device=img.device, dtype=torch.float32)
    minmax = torch.empty(B * Cout, 2, device=img.device, dtype=torch.float32) if want_minmax else None
    with _Launch(img, seg) as st:
        _lib.check(_lib.load().swn_sigmoid_mask(_ptr(img), Cimg, _ptr(seg), _ptr(images2), _ptr(seg_map), _ptr(masked),
                                                _ptr(minmax), B, Cout, H, W, st), "swn_sigmoid_mask")
    _count("sigmoid_mask_mm" if want_minmax else "sigmoid_mask")
    return (images2 if images2 is not None else img), seg_map, masked, minmax


def ensure_2ch(x):
    """[B,1,H,W] fp32 -> [B,2,H,W]: channel 1 = sqrt(|channel 0|) (ST_Inference_Pipline.py:32-37)."""
    B, C, H, W = x.shape
    out = torch.empty(B, 2, H, W, device=x.device, dtype=torch.float32)
    with _Launch(x) as st:
        _lib.check(_lib.load().swn_ensure_2ch(_ptr(x), _ptr(out), B, H * W, st), "swn_ensure_2ch")
    _count("ensure_2ch")
    return out


def normalize(x, minmax, inverse, threshold=0.01, eps=1e-6):
    B, C, H, W = x.shape
    if tuple(minmax.shape) != (B * C, 2) or minmax.dtype != torch.float32 or x.dtype != torch.float32 or not x.is_contiguous() \
            or not minmax.is_contiguous():
        raise RuntimeError(f"swinwnet_b200: normalize needs contiguous fp32 x [B,C,H,W] and minmax [{B * C},2] "
                           f"(got {tuple(x.shape)} / {tuple(minmax.shape)})")
    out = torch.empty_like(x)
    with _Launch(x, minmax) as st:
        _lib.check(_lib.load().swn_normalize(_ptr(x), _ptr(minmax), _ptr(out), B * C, H, W, threshold, eps, int(inverse), st),
                   "swn_normalize")
    _count("normalize")
    return out


def dspace_histogram(img, bin_of_pixel, n_bins):
    """img [B, C, H, W] fp32 CUDA (channel 0 is used), bin_of_pixel int32 [H*W] -> [B, n_bins] fp32 (one launch)."""
    B, C, H, W = img.shape
    img = img.float().contiguous()
    out = torch.empty(B, n_bins, device=img.device, dtype=torch.float32)
    with _Launch(img, bin_of_pixel) as st:
        _lib.check(_lib.load().swn_dspace_histogram(_ptr(img), C * H * W, _ptr(bin_of_pixel), B, H * W, n_bins, _ptr(out), st),
                   "swn_dspace_histogram")
    _count("dspace_histogram")
    return out


def adamw_multi(table, chunks, n_chunks, lr, beta1, beta2, eps, weight_decay, grad_scale=1.0):
    """one launch of torch.optim.AdamW semantics over every parameter described by the device table (train.FusedAdamW)."""
    with _Launch(table, chunks) as st:
        _lib.check(_lib.load().swn_adamw_multi(_ptr(table), _ptr(chunks), n_chunks, lr, beta1, beta2, eps, weight_decay,
                                               grad_scale, st), "swn_adamw_multi")
    _count("adamw_multi")


def grad_bucket_copy(table, chunks, n_chunks, flat, unpack, scale=1.0):
    with _Launch(table, chunks, flat) as st:
        _lib.check(_lib.load().swn_grad_bucket_copy(_ptr(table), _ptr(chunks), n_chunks, _ptr(flat), int(unpack), scale, st),
                   "swn_grad_bucket_copy")
    _count("grad_bucket_copy")
