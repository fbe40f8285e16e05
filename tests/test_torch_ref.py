"""CPU tests of the backward-side restatements (``torch_ref.py``): every leaf operator against the pinned oracle on
seeded inputs (values to 1e-5), so that the gradients the autograd boundary produces are gradients of the reference's
function.  The forward kernels themselves are covered by the GPU tests."""
import pytest
import torch

import swinwnet_b200 as S
from swinwnet_b200 import torch_ref as T
from oracle import swinwnet_oracle as O


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def close(a, b, tol=2e-5):
    assert a.shape == b.shape
    assert (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


def _block_sd(C, nH):
    shapes = {"norm1.weight": (C,), "norm1.bias": (C,), "attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
              "attn.relative_position_bias_table": (81, nH), "attn.proj.weight": (C, C), "attn.proj.bias": (C,),
              "norm2.weight": (C,), "norm2.bias": (C,), "mlp.0.weight": (4 * C, C), "mlp.0.bias": (4 * C,),
              "mlp.3.weight": (C, 4 * C), "mlp.3.bias": (C,)}
    sd = {k: rnd(*s, seed=10 + i) * ((s[-1] ** -0.5) if len(s) == 2 and "table" not in k else 0.2) for i, (k, s) in enumerate(shapes.items())}
    sd["norm1.weight"] += 1.0
    sd["norm2.weight"] += 1.0
    return sd


@pytest.mark.parametrize("C,nH,H,W", [(12, 3, 7, 11), (48, 6, 10, 5), (96, 3, 6, 6)])
def test_block_matches_oracle(C, nH, H, W):
    sd = _block_sd(C, nH)
    x = rnd(2, H * W, C, seed=1)
    close(T.swin_block(x, (H, W), nH, *[sd[k] for k in sd]), O.swin_block(sd, "", x, (H, W), nH, 0))


@pytest.mark.parametrize("H,W", [(9, 13), (8, 6)])
def test_merge_expand_linear_match_oracle(H, W):
    C = 24
    x = rnd(2, H * W, C, seed=1)
    sd = {"reduction.weight": rnd(2 * C, 4 * C, seed=2, scale=0.1), "norm.weight": 1 + 0.1 * rnd(4 * C, seed=3), "norm.bias": 0.1 * rnd(4 * C, seed=4)}
    ref, _ = O.patch_merging(sd, "", x, (H, W))
    close(T.patch_merging(x, (H, W), sd["reduction.weight"], sd["norm.weight"], sd["norm.bias"]), ref)
    sd = {"expand.weight": rnd(2 * C, C, seed=2, scale=0.2), "norm.weight": 1 + 0.1 * rnd(C // 2, seed=3), "norm.bias": 0.1 * rnd(C // 2, seed=4)}
    ref, res = O.patch_expanding(sd, "", x, (H, W))
    close(T.patch_expanding(x, (H, W), None, sd["expand.weight"], sd["norm.weight"], sd["norm.bias"]), ref)
    tgt = (2 * H - 1, 2 * W - 1)
    close(T.patch_expanding(x, (H, W), tgt, sd["expand.weight"], sd["norm.weight"], sd["norm.bias"]), O.crop_tokens(ref, res, tgt))


@pytest.mark.parametrize("Cin,H,W,s", [(2, 40, 60, 1), (1, 35, 51, 1), (2, 80, 120, 2)])
def test_patch_embed_matches_oracle(Cin, H, W, s):
    x = rnd(2, Cin, H, W, seed=1) * 3
    sd = {"proj.weight": rnd(48, Cin, 2, 2, seed=2, scale=0.5), "proj.bias": rnd(48, seed=3, scale=0.1),
          "norm.weight": 1 + 0.1 * rnd(48, seed=4), "norm.bias": 0.1 * rnd(48, seed=5)}
    ref, _ = O.patch_embed(sd, "", x, s)
    close(T.patch_embed(x, sd["proj.weight"], sd["proj.bias"], sd["norm.weight"], sd["norm.bias"], scale=s), ref)


def test_heads_and_cross_attention_match_oracle():
    B, Hq, Wq = 2, 9, 14
    x = rnd(B, Hq * Wq, 48, seed=1)
    sd = {"seg_head.0.weight": rnd(24, 48, 3, 3, seed=2, scale=0.05), "seg_head.0.bias": rnd(24, seed=3, scale=0.1),
          "seg_head.2.weight": rnd(1, 24, 1, 1, seed=4, scale=0.2), "seg_head.2.bias": rnd(1, seed=5, scale=0.1)}
    for scale in (1, 2):
        res = (Hq * 2 * scale, Wq * 2 * scale)
        close(T.segmentation_head(x, res, scale, *[sd[k] for k in sd]), O.segmentation_head(sd, "", x, res, scale), 1e-4)
    y = rnd(B, 12 * 20, 12, seed=1)
    w1, b1, w2, b2 = rnd(12, 12, 3, 3, seed=2, scale=0.1), rnd(12, seed=3, scale=0.1), rnd(2, 12, 1, 1, seed=4, scale=0.3), rnd(2, seed=5, scale=0.1)
    h = O.gelu_erf(O.conv3x3_nhwc(y.view(B, 12, 20, 12), w1, b1))
    ref = (h @ w2.view(2, 12).t() + b2).permute(0, 3, 1, 2)[:, :, :10, :18]
    close(T.recon_tail(y, (12, 20), (10, 18), w1, b1, w2, b2), ref, 1e-4)
    C = 192
    q, kv = rnd(2, 50, C, seed=1), rnd(2, 70, C, seed=2)
    sd = {"norm_q.weight": 1 + 0.1 * rnd(C, seed=3), "norm_q.bias": 0.1 * rnd(C, seed=4), "norm_kv.weight": 1 + 0.1 * rnd(C, seed=5),
          "norm_kv.bias": 0.1 * rnd(C, seed=6), "attn.in_proj_weight": rnd(3 * C, C, seed=7, scale=C ** -0.5),
          "attn.in_proj_bias": rnd(3 * C, seed=8, scale=0.1), "attn.out_proj.weight": rnd(C, C, seed=9, scale=C ** -0.5),
          "attn.out_proj.bias": rnd(C, seed=10, scale=0.1), "gamma": torch.tensor([0.4])}
    ref = O.cross_attention_block(sd, "", q, kv, 3)
    got = T.cross_attention_block(q, kv, 3, sd["attn.in_proj_weight"], sd["attn.in_proj_bias"], sd["attn.out_proj.weight"],
                                  sd["attn.out_proj.bias"], sd["norm_q.weight"], sd["norm_q.bias"], sd["norm_kv.weight"],
                                  sd["norm_kv.bias"], sd["gamma"])
    close(got, ref)


def test_kernel_op_backward_is_the_gradient_of_the_restatement():
    """KernelOp's backward (recompute + autograd.grad, honouring requires_grad flags) on CPU with a stand-in "kernel":
    gradients equal those of the plain torch function; frozen inputs get None."""
    from swinwnet_b200.autograd import op
    x = rnd(5, 4, seed=1).requires_grad_(True)
    w = rnd(3, 4, seed=2).requires_grad_(True)
    b = rnd(3, seed=3)                                       # frozen
    f = lambda x_, w_, b_: torch.nn.functional.gelu(torch.nn.functional.linear(x_, w_, b_))
    y = op(lambda *t: f(*t).detach(), f, x, w, b)
    y.square().sum().backward()
    gx, gw = x.grad.clone(), w.grad.clone()
    x.grad = w.grad = None
    f(x, w, b).square().sum().backward()
    close(gx, x.grad)
    close(gw, w.grad)
    assert b.grad is None
