"""GPU parity tests added in round 2 (VERDICT r1 "Next round" item 1 and the ADVICE findings):

* persistent kernels with >= 20 tiles per CTA (mbarrier parities, weight rings and TMEM buffers wrap dozens of times),
  against the fp32 oracle evaluated on the same device (plain torch, TF32 off);
* range / conditioning edge cases of the 16-bit operand path (large residual stream, |mean| >> std LayerNorm rows);
* the single-branch models at the full 250x480 geometry (BASELINE config 3);
* boundary behaviour: odd image sizes raise like the reference, micro-batched large batches, non-current device.
"""
import pytest
import torch

import swinwnet_b200 as S
from swinwnet_b200 import ops, packing
from oracle import swinwnet_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-2
OPD = S.ops.operand_dtype()
D2 = [2, 2, 2, 2]


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    """the on-device oracle must be true fp32 (no TF32 in matmuls / convolutions)"""
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b
    torch.cuda.empty_cache()


def relerr(a, b):
    a, b = a.detach().float(), b.detach().float().to(a.device)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.isfinite(a).all(), "non-finite values in kernel output"
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=DEV) * scale


def _block_sd(C, nH, seed=10):
    shapes = {"norm1.weight": (C,), "norm1.bias": (C,), "attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
              "attn.relative_position_bias_table": (81, nH), "attn.proj.weight": (C, C), "attn.proj.bias": (C,),
              "norm2.weight": (C,), "norm2.bias": (C,), "mlp.0.weight": (4 * C, C), "mlp.0.bias": (4 * C,),
              "mlp.3.weight": (C, 4 * C), "mlp.3.bias": (C,)}
    sd = {k: rnd(*s, seed=seed + i) * ((s[-1] ** -0.5) if "weight" in k and len(s) == 2 else 0.2)
          for i, (k, s) in enumerate(shapes.items())}
    sd["norm1.weight"] += 1.0
    sd["norm2.weight"] += 1.0
    return sd


def _sms():
    return torch.cuda.get_device_properties(0).multi_processor_count


# ---------------------------------------------------------------------------------------------
# deep tiles: >= 20 tiles per persistent CTA
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [48, 96, 192])
def test_mlp_deep_tiles(C):
    M = 128 * _sms() * 22 + 77           # 22 full tiles per CTA + a ragged tail
    x = rnd(M, C, seed=1)
    W1, b1 = rnd(4 * C, C, seed=2, scale=C ** -0.5), rnd(4 * C, seed=3, scale=0.2)
    W2, b2 = rnd(C, 4 * C, seed=4, scale=(4 * C) ** -0.5), rnd(C, seed=5, scale=0.1)
    lw, lb = 1 + 0.1 * rnd(C, seed=6), 0.1 * rnd(C, seed=7)
    ref = x + O.linear(O.gelu_erf(O.linear(O.layer_norm(x, lw, lb), W1, b1)), W2, b2)
    HC, TR = ops.mlp_config(C)
    Wp, b2p = packing.pack_mlp(W1, W2, b2, HC, TR)
    out = torch.full_like(x, float("nan"))
    ops.mlp(x, out, M, C, lw, lb, Wp, b1, b2p)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL
    # every tile individually: the worst 128-row tile must be as good as the global figure suggests
    err = (out - ref).abs().amax(dim=1)
    worst = err.view(-1)[: (M // 128) * 128].view(-1, 128).amax(dim=1)
    assert worst.max().item() <= TOL * ref.abs().max().item()


@pytest.mark.parametrize("K,N,a_mode,e_mode", [(96, 48, "f32", "f32"), (48, 144, "ln", "op"), (192, 96, "f32", "f32"),
                                               (192, 576, "ln", "op"), (96, 96, "op", "res")])
def test_rowgemm_deep_tiles(K, N, a_mode, e_mode):
    M = 128 * _sms() * 21 + 5
    W, b = rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3, scale=0.1)
    lw, lb = 1 + 0.1 * rnd(K, seed=4), 0.1 * rnd(K, seed=5)
    nv = packing.choose_chunk(N, 128 if K <= 192 else 256)
    Wp, bp, NT, nch = packing.pack_rowgemm(W, b, nv)
    if a_mode == "op":
        A = rnd(M, K, seed=1).to(OPD)
        res = rnd(M, N, seed=6)
        ref = res + O.linear(A.float(), W, b)
        out = torch.full((M, N), float("nan"), device=DEV)
        ops.rowgemm(A=A, a_mode=ops.A_BF16, M=M, K=K, lda=K, Wp=Wp, NT=NT, nchunks=nch, n_valid=nv, e_mode=ops.E_F32,
                    bias=bp, out=out, ldo=N, res=res, ldres=N)
    else:
        A = rnd(M, K, seed=1) * 1.5 + 0.3
        ref = O.linear(O.layer_norm(A, lw, lb) if a_mode == "ln" else A, W, b)
        out = torch.empty(M, N, device=DEV, dtype=OPD if e_mode == "op" else torch.float32)
        ops.rowgemm(A=A, a_mode=ops.A_F32_LN if a_mode == "ln" else ops.A_F32, M=M, K=K, lda=K, ln_w=lw, ln_b=lb, Wp=Wp, NT=NT,
                    nchunks=nch, n_valid=nv, e_mode=ops.E_BF16 if e_mode == "op" else ops.E_F32, bias=bp, out=out, ldo=N)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL


@pytest.mark.parametrize("C,nH,B,H,W", [(48, 3, 16, 125, 240), (96, 3, 16, 125, 240), (96, 6, 64, 63, 120),
                                        (24, 3, 8, 250, 480), (12, 3, 3, 500, 960)])
def test_fused_block_deep_tiles(C, nH, B, H, W):
    """the fused W-MSA / whole-block kernels at the model's own geometries with >= 20 tiles per persistent CTA"""
    x = rnd(B, H * W, C, seed=1) * 1.5 + 0.2
    sd = _block_sd(C, nH)
    whole = C < 96
    if whole:
        ref = O.swin_block(sd, "", x, (H, W), nH, 0)
        Wpk, fpk = packing.pack_fused_block(sd["norm1.weight"], sd["norm1.bias"], sd["attn.qkv.weight"], sd["attn.qkv.bias"],
                                            sd["attn.relative_position_bias_table"], sd["attn.proj.weight"], sd["attn.proj.bias"],
                                            sd["norm2.weight"], sd["norm2.bias"], sd["mlp.0.weight"], sd["mlp.0.bias"],
                                            sd["mlp.3.weight"], sd["mlp.3.bias"], nH)
    else:
        ref = x + O.window_attention(sd, "attn.", O.layer_norm(x, sd["norm1.weight"], sd["norm1.bias"]), (H, W), nH, 0)
        Wpk, fpk = packing.pack_fused_attn_stream(sd["norm1.weight"], sd["norm1.bias"], sd["attn.qkv.weight"], sd["attn.qkv.bias"],
                                                  sd["attn.relative_position_bias_table"], sd["attn.proj.weight"],
                                                  sd["attn.proj.bias"], nH)
    out = torch.full_like(x, float("nan"))
    ops.swin_block_fused(x, out, B, H, W, C, nH, 1e-5, Wpk, fpk, whole)
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL / 4
    # second launch into the same buffers must be bit-identical (no state leaks between launches)
    out2 = torch.full_like(x, float("nan"))
    ops.swin_block_fused(x, out2, B, H, W, C, nH, 1e-5, Wpk, fpk, whole)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


@pytest.mark.parametrize("C,nH,B,H,W", [(192, 6, 64, 63, 120), (384, 12, 64, 32, 60), (384, 24, 64, 16, 30), (192, 12, 64, 32, 60)])
def test_wide_blocks_model_geometry(C, nH, B, H, W):
    """C >= 192 blocks (decoder stage 0/1, encoder stage 2/3, bottleneck) at batch 64 through the module lowering"""
    sd = _block_sd(C, nH)
    blk = S.model.SwinTransformerBlock(C, nH).to(DEV)
    blk.load_state_dict({**sd, "attn.relative_position_index": O.rel_pos_index()}, strict=True)
    x = rnd(B, H * W, C, seed=1) * 1.5 + 0.2
    ref = O.swin_block(sd, "", x, (H, W), nH, 0)
    with torch.no_grad():
        out = blk(x, (H, W))
    torch.cuda.synchronize()
    assert relerr(out, ref) <= TOL / 2


# ---------------------------------------------------------------------------------------------
# range / conditioning of the 16-bit operand path
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scale", [1e3, 1e6])
def test_a_f32_prologue_large_residual_stream(scale):
    """PatchExpanding.expand / decoder linears convert the RAW fp32 residual stream to the 16-bit operand type
    (model.py PatchExpanding.run / SwinDecoder.forward).  x1e3: inside fp16 range, full parity.  x1e6: beyond it — the
    conversion saturates (fp16 build) instead of producing inf/NaN, and the result stays finite."""
    C, H, W, B = 96, 20, 30, 2
    x = rnd(B, H * W, C, seed=1) * scale
    sd = {"expand.weight": rnd(2 * C, C, seed=2, scale=C ** -0.5), "norm.weight": 1 + 0.1 * rnd(C // 2, seed=3),
          "norm.bias": 0.1 * rnd(C // 2, seed=4)}
    m = S.model.PatchExpanding(C).to(DEV)
    m.load_state_dict(sd)
    with torch.no_grad():
        out, _ = m.run(x, (H, W))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    if scale <= 1e3 or OPD == torch.bfloat16:
        ref, _ = O.patch_expanding(sd, "", x, (H, W))
        assert relerr(out, ref) <= TOL
    # decoder-style linear on the raw stream
    Wl, bl = rnd(C // 2, C, seed=5, scale=C ** -0.5), rnd(C // 2, seed=6, scale=0.1)
    Wp, bp, NT, nch = packing.pack_rowgemm(Wl, bl, C // 2)
    y = torch.empty(B * H * W, C // 2, device=DEV)
    ops.rowgemm(A=x.view(-1, C), a_mode=ops.A_F32, M=B * H * W, K=C, lda=C, Wp=Wp, NT=NT, nchunks=nch, n_valid=C // 2,
                e_mode=ops.E_F32, bias=bp, out=y, ldo=C // 2)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    if scale <= 1e3 or OPD == torch.bfloat16:
        assert relerr(y, O.linear(x.view(-1, C), Wl, bl)) <= TOL


@pytest.mark.parametrize("C", [12, 48, 96, 192, 384])
def test_layernorm_rows_with_large_mean(C):
    """|mean| >> std rows (1e4 +- 1): the kernels use shifted one-pass moments (shift = first element of the row), which
    must not cancel; compared against torch's two-pass fp32 LayerNorm through a whole block / MLP."""
    M = 3000
    x = rnd(M, C, seed=1) + 1.0e4
    W1, b1 = rnd(4 * C, C, seed=2, scale=C ** -0.5), rnd(4 * C, seed=3, scale=0.2)
    W2, b2 = rnd(C, 4 * C, seed=4, scale=(4 * C) ** -0.5), rnd(C, seed=5, scale=0.1)
    lw, lb = 1 + 0.1 * rnd(C, seed=6), 0.1 * rnd(C, seed=7)
    branch = O.linear(O.gelu_erf(O.linear(O.layer_norm(x, lw, lb), W1, b1)), W2, b2)
    HC, TR = ops.mlp_config(C)
    Wp, b2p = packing.pack_mlp(W1, W2, b2, HC, TR)
    out = torch.empty_like(x)
    ops.mlp(x, out, M, C, lw, lb, Wp, b1, b2p)
    torch.cuda.synchronize()
    got = out.double() - x.double()          # the residual (1e4) would hide the branch: compare the branch itself
    # the fp32 residual add rounds at ulp(1e4) ~ 1e-3: that is the floor of this comparison
    assert (got.float() - branch).abs().max().item() <= TOL * branch.abs().max().item() + 2e-3
    if C in (12, 48):
        sd = _block_sd(C, 3)
        xb = (rnd(2, 10 * 15, C, seed=9) + 1.0e4)
        ref = O.swin_block(sd, "", xb, (10, 15), 3, 0)
        blk = S.model.SwinTransformerBlock(C, 3).to(DEV)
        blk.load_state_dict({**sd, "attn.relative_position_index": O.rel_pos_index()}, strict=True)
        with torch.no_grad():
            o = blk(xb, (10, 15))
        assert ((o - xb) - (ref - xb)).abs().max().item() <= TOL * (ref - xb).abs().max().item() + 4e-3


@pytest.mark.parametrize("C", [48, 96, 192])
@pytest.mark.parametrize("bias_scale", [0.0, 1e5])
def test_gelu_large_preactivations(C, bias_scale):
    """hidden pre-activations of a few thousand (in fp16 range: full parity through the clamped tanh-form GELU) and, with
    bias_scale = 1e5, beyond it: the GELU output saturates at +-65504 (fp16 build) and the result stays finite."""
    M = 700
    x = rnd(M, C, seed=1)
    W1, b1 = rnd(4 * C, C, seed=2, scale=300.0), rnd(4 * C, seed=3) * max(bias_scale, 1.0)
    W2, b2 = rnd(C, 4 * C, seed=4, scale=1e-4), rnd(C, seed=5, scale=0.1)
    lw, lb = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    HC, TR = ops.mlp_config(C)
    Wp, b2p = packing.pack_mlp(W1, W2, b2, HC, TR)
    out = torch.empty_like(x)
    ops.mlp(x, out, M, C, lw, lb, Wp, b1, b2p)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    if bias_scale == 0.0 or OPD == torch.bfloat16:
        ref = x + O.linear(O.gelu_erf(O.linear(O.layer_norm(x, lw, lb), W1, b1)), W2, b2)
        assert relerr(out, ref) <= TOL


# ---------------------------------------------------------------------------------------------
# models at full geometry, boundary behaviour
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["unet", "unetsr"])
def test_single_branch_models_full_geometry(manifest, name):
    """BASELINE config 3 geometry (250x480), B = 2, against the fp32 oracle evaluated on the device"""
    sd = O.make_state_dict(manifest[name], seed=1)
    cls, fn = (S.SwinUNet, O.swin_unet) if name == "unet" else (S.SwinUNetSR, O.swin_unet_sr)
    m = cls(depths=D2)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    x = O.normalize_piecewise(O.synthetic_diffractions(2, seed=61, two_channel=False))[0].to(DEV)
    with torch.no_grad():
        y = m(x)
    ref = fn({k: v.to(DEV) for k, v in sd.items()}, x)
    assert y.shape == ref.shape == ((2, 1, 250, 480) if name == "unet" else (2, 1, 500, 960))
    assert relerr(y, ref) <= TOL
    # and against the same oracle on the CPU for sample 0 (ties the on-device oracle to the pinned one)
    ref0 = fn(sd, x[:1].cpu())
    assert relerr(ref[:1].cpu(), ref0) <= 1e-4


def test_odd_image_size_raises_like_the_reference(manifest):
    """segment_1 returns the padded size for odd H / W; the reference's `images * seg_map` then fails to broadcast
    (ST_Inference_Pipline.py:96).  The drop-in must raise as well instead of masking misaligned rows."""
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(O.make_state_dict(manifest["wnet_em"], seed=1), strict=True)
    inf = S.SwinWNetInference(m, DEV)
    with pytest.raises(RuntimeError, match="must match the size"):
        inf(O.synthetic_diffractions(1, seed=2, H=35, W=51, two_channel=False).to(DEV))


def test_micro_batched_large_batch_matches_small_batches(manifest):
    """B > max_batch: chunks written into outputs allocated once; every sample equals its solo result bit for bit"""
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(O.make_state_dict(manifest["wnet_em"], seed=1), strict=True)
    x = O.synthetic_diffractions(11, seed=5, H=40, W=60, two_channel=False).to(DEV)
    big = S.SwinWNetInference(m, DEV, max_batch=4)
    out = big(x).clone()
    stages = {k: getattr(big, k).clone() for k in big._STAGES}
    assert out.shape == (11, 2, 80, 120)
    one = S.SwinWNetInference(m, DEV, max_batch=64)
    for i in (0, 3, 4, 10):
        solo = one(x[i:i + 1])
        assert torch.equal(out[i:i + 1], solo)
        for k in big._STAGES:
            assert torch.equal(stages[k][i:i + 1], getattr(one, k)), k
    # host path: B = 0 and ragged chunks
    o0 = big.run_host(torch.zeros(0, 1, 40, 60))
    assert o0.shape == (0, 2, 80, 120)
    oh = big.run_host(x.cpu().pin_memory(), chunk=3)
    big.host_done.synchronize()
    assert torch.equal(oh, out.cpu())


def test_batch_1024_full_geometry_micro_batched(manifest):
    """BASELINE config 4: a batch far above max_batch at the dataset geometry runs in a bounded workspace and every
    micro-batch boundary sample equals its solo result"""
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(O.make_state_dict(manifest["wnet_em"], seed=1), strict=True)
    inf = S.SwinWNetInference(m, DEV, max_batch=64)
    base = O.synthetic_diffractions(4, seed=77, two_channel=False)
    B = 1024
    x = (base.repeat(B // 4, 1, 1, 1) * (1.0 + 0.001 * torch.arange(B).view(B, 1, 1, 1))).to(DEV)
    torch.cuda.reset_peak_memory_stats()
    out = inf(x)
    torch.cuda.synchronize()
    assert out.shape == (B, 2, 500, 960) and torch.isfinite(out).all()
    peak = torch.cuda.max_memory_allocated() / 2 ** 30
    print(f"peak memory at B={B}: {peak:.1f} GiB")
    keep = {i: out[i:i + 1].clone() for i in (0, 63, 64, 1023)}
    del out
    inf._reset_outputs()
    torch.cuda.empty_cache()
    for i, ref in keep.items():
        assert torch.equal(inf(x[i:i + 1]), ref), i


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_runs_on_a_non_current_device(manifest):
    """SwinWNetInference(model, 'cuda:1') while the current device is 0: launches must go to GPU 1"""
    torch.cuda.set_device(0)
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(O.make_state_dict(manifest["wnet_em"], seed=1), strict=True)
    x = O.synthetic_diffractions(2, seed=5, H=40, W=60, two_channel=False)
    a = S.SwinWNetInference(m, "cuda:1")(x.to("cuda:1")).cpu()
    assert torch.cuda.current_device() == 0
    b = S.SwinWNetInference(m, "cuda:0")(x.to("cuda:0")).cpu()
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match="different devices"):
        ops.normalize(x.to("cuda:0"), torch.zeros(2, 2, device="cuda:1"), inverse=False)


# ---------------------------------------------------------------------------------------------
# compute-sanitizer is closed on this GPU pool (profiles/r2_compute_sanitizer_closed.txt), so the memcheck / racecheck
# runs SURVEY.md §5 asks for are replaced by what can be checked from the outside: guard bands around every output
# (out-of-bounds writes), NaN-prefilled outputs (unwritten elements) and bit-identical repeated launches of the
# persistent kernels with many tiles per CTA (a shared-memory / mbarrier race shows up as run-to-run differences).
# ---------------------------------------------------------------------------------------------
def _guarded(shape, dtype=torch.float32, pad=4096):
    n = 1
    for s in shape:
        n *= s
    buf = torch.full((n + 2 * pad,), 12345.0, device=DEV, dtype=dtype)
    view = buf[pad:pad + n].view(*shape)
    view.fill_(float("nan"))
    return buf, view, pad


def _guards_intact(buf, pad):
    return bool((buf[:pad] == 12345.0).all() and (buf[-pad:] == 12345.0).all())


@pytest.mark.parametrize("C,nH,B,H,W", [(12, 3, 2, 123, 241), (24, 3, 2, 63, 121), (48, 6, 3, 62, 99), (96, 3, 3, 62, 99), (96, 6, 5, 63, 120)])
def test_fused_kernels_guard_bands_and_repeatability(C, nH, B, H, W):
    x = rnd(B, H * W, C, seed=1)
    sd = _block_sd(C, nH)
    whole = C < 96
    pk = (packing.pack_fused_block(*[sd[k] for k in sd], nH) if whole else
          packing.pack_fused_attn_stream(*[sd[k] for k in list(sd)[:7]], nH))
    ref = None
    for rep in range(6):
        buf, out, pad = _guarded((B, H * W, C))
        ops.swin_block_fused(x, out, B, H, W, C, nH, 1e-5, pk[0], pk[1], whole)
        torch.cuda.synchronize()
        assert _guards_intact(buf, pad) and torch.isfinite(out).all()
        ref = out.clone() if ref is None else ref
        assert torch.equal(out, ref), f"launch {rep} differs from launch 0"


@pytest.mark.parametrize("C", [48, 96, 192, 384])
def test_mlp_and_rowgemm_guard_bands_and_repeatability(C):
    M = 128 * _sms() * 3 + 37
    x = rnd(M, C, seed=1)
    W1, b1 = rnd(4 * C, C, seed=2, scale=C ** -0.5), rnd(4 * C, seed=3, scale=0.2)
    W2, b2 = rnd(C, 4 * C, seed=4, scale=(4 * C) ** -0.5), rnd(C, seed=5, scale=0.1)
    lw, lb = 1 + 0.1 * rnd(C, seed=6), 0.1 * rnd(C, seed=7)
    HC, TR = ops.mlp_config(C)
    Wp, b2p = packing.pack_mlp(W1, W2, b2, HC, TR)
    Wq, bq = rnd(3 * C, C, seed=8, scale=C ** -0.5), rnd(3 * C, seed=9, scale=0.1)
    nv = packing.choose_chunk(3 * C, 128 if C <= 192 else 256)
    Wqp, bqp, NT, nch = packing.pack_rowgemm(Wq, bq, nv)
    ref_m = ref_q = None
    for rep in range(5):
        buf, out, pad = _guarded((M, C))
        ops.mlp(x, out, M, C, lw, lb, Wp, b1, b2p)
        bq2, qkv, pad2 = _guarded((M, 3 * C), dtype=OPD)
        ops.rowgemm(A=x, a_mode=ops.A_F32_LN, M=M, K=C, lda=C, ln_w=lw, ln_b=lb, Wp=Wqp, NT=NT, nchunks=nch, n_valid=nv,
                    e_mode=ops.E_BF16, bias=bqp, out=qkv, ldo=3 * C)
        torch.cuda.synchronize()
        assert _guards_intact(buf, pad) and _guards_intact(bq2, pad2)
        assert torch.isfinite(out).all() and torch.isfinite(qkv.float()).all()
        ref_m, ref_q = (out.clone(), qkv.clone()) if ref_m is None else (ref_m, ref_q)
        assert torch.equal(out, ref_m) and torch.equal(qkv, ref_q), f"launch {rep} differs from launch 0"
