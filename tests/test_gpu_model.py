"""GPU parity tests, model level: the drop-in modules (CUDA path, through the C ABI) against
(1) the committed golden vectors produced by the unmodified reference and (2) the oracle on seeded inputs.
north_star tolerances: logits / upscaled images max-norm relative error <= 2e-2 (bf16 tensor-core operands);
PSNR within 0.05 dB; size-independent properties at the full 250x480 geometry."""
import math

import pytest
import torch

import swinwnet_b200 as S
from oracle import swinwnet_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16 = S.ops.operand_dtype() == torch.bfloat16     # non-default build variant: characterisation bounds (see test_gpu_gates.py)
TOL = 1e-1 if BF16 else 2e-2
D2 = [2, 2, 2, 2]


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.isfinite(a).all()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


def psnr(a, t):
    a, t = a.clamp(0, 1), t.clamp(0, 1)
    return 10.0 * math.log10(1.0 / max(((a - t) ** 2).mean().item(), 1e-12))


@pytest.fixture(scope="module")
def wnet_em(manifest):
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(O.make_state_dict(manifest["wnet_em"], seed=1), strict=True)
    return m.to(DEV).eval()


def test_case_A_pipeline_vs_reference_golden(wnet_em, golden):
    x = O.synthetic_diffractions(2, seed=3, H=40, W=60, two_channel=False)
    inf = S.SwinWNetInference(wnet_em, DEV)
    out = inf(x.to(DEV))
    with torch.no_grad():
        seg, skips = wnet_em.segment_1(inf.images)
    assert relerr(seg, golden["A_seg_lr_logits"]) <= TOL
    for i, s in enumerate(skips):
        assert relerr(s, golden[f"A_skip{i}"]) <= TOL, i
    for k in ("seg_map_lr", "norm", "upscaled_norm", "upscaled_denorm", "seg_map_hr", "images_masked_hr"):
        assert relerr(getattr(inf, k), golden["A_" + k]) <= TOL, k
    assert out is inf.images_masked_hr
    # PSNR gate (tests.py:349-357 protocol: clamp(0,1), data_range 1) against the normalised HR reference output
    t = golden["A_upscaled_norm"]
    assert psnr(inf.upscaled_norm.cpu(), t) > 40.0


def test_case_B_diffraction_only(manifest, golden):
    m = S.SwinWNet(error_matrix=False, depths=D2)
    m.load_state_dict(O.make_state_dict(manifest["wnet"], seed=1), strict=True)
    m = m.to(DEV).eval()
    x = O.synthetic_diffractions(1, seed=4, H=36, W=50, two_channel=False)
    inf = S.SwinWNetInference(m, DEV)
    inf(x.to(DEV), two_channel=False)
    assert relerr(inf.seg_lr_logits, golden["B_seg_lr_logits"]) <= TOL
    assert relerr(inf.upscaled_norm, golden["B_upscaled_norm"]) <= TOL
    assert relerr(inf.seg_hr_logits, golden["B_seg_hr_logits"]) <= TOL
    # odd image size: patch-embed pad path, head returns the padded size (reference behaviour)
    x = O.synthetic_diffractions(1, seed=8, H=35, W=51, two_channel=False)
    with torch.no_grad():
        seg, _ = m.segment_1((x / 100.0).to(DEV))
    assert relerr(seg, golden["B2_seg_lr_logits"]) <= TOL


def test_case_C_single_branch_models(manifest, golden):
    x = O.synthetic_diffractions(1, seed=5, H=30, W=44, two_channel=False)
    xn, _ = O.normalize_piecewise(x)
    for name, cls in (("unet", S.SwinUNet), ("unetsr", S.SwinUNetSR)):
        m = cls(depths=D2)
        m.load_state_dict(O.make_state_dict(manifest[name], seed=1), strict=True)
        with torch.no_grad():
            y = m.to(DEV).eval()(xn.to(DEV))
        assert relerr(y, golden["C_" + name]) <= TOL, name


def test_case_D_even_training_step_geometry(wnet_em, golden):
    x = O.synthetic_diffractions(1, seed=6, H=40, W=60)
    xn, _ = O.normalize_piecewise(x)
    with torch.no_grad():
        _, sk = wnet_em.segment_1(xn.to(DEV))
        up, _ = wnet_em.upscale(golden["D_lr_input"].to(DEV), sk)
    assert relerr(up, golden["D_upscaled"]) <= TOL


def test_case_E_full_geometry_vs_reference_golden(wnet_em, golden):
    x = O.synthetic_diffractions(1, seed=7, two_channel=False)
    inf = S.SwinWNetInference(wnet_em, DEV)
    inf(x.to(DEV))
    assert relerr(inf.seg_lr_logits[:, :, ::5, ::5], golden["E_seg_lr_logits_s5"]) <= TOL
    assert relerr(inf.upscaled_norm[:, :, ::5, ::5], golden["E_upscaled_norm_s5"]) <= TOL
    assert relerr(inf.seg_map_hr[:, :, ::5, ::5], golden["E_seg_map_hr_s5"]) <= TOL
    assert relerr(inf.images_masked_hr[:, :, ::5, ::5], golden["E_images_masked_hr_s5"]) <= TOL
    # binarised masks at 0.5 (tests.py:12-16)
    a = (inf.seg_map_hr[:, :, ::5, ::5].cpu() >= 0.5)
    b = (golden["E_seg_map_hr_s5"] >= 0.5)
    agree = (a == b).float().mean().item()
    print("mask agreement (random-init weights, HR):", agree)
    assert agree >= 0.99


def test_oracle_parity_random_seed(manifest):
    """fresh seeds (not in the golden set) against the oracle executed here on CPU."""
    sd = O.make_state_dict(manifest["wnet_em"], seed=5)
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    x = O.synthetic_diffractions(3, seed=11, H=50, W=70)
    ref = O.st_pipeline(sd, x)
    inf = S.SwinWNetInference(m, DEV, max_batch=2)       # exercises micro-batching (2 + 1)
    inf(x.to(DEV))
    for k in ("seg_lr_logits", "upscaled_norm", "seg_hr_logits", "images_masked_hr"):
        assert relerr(getattr(inf, k), ref[k]) <= TOL, k


@pytest.mark.parametrize("H,W", [(22, 18), (36, 50), (10, 130), (62, 44)])
def test_ragged_and_tiny_geometries_against_oracle(manifest, H, W):
    """edge geometries: grids that are no multiple of the window (5), the patch (2) or the merge (2) at some scale, and
    images so small that the deepest stage is a single, mostly padded window — every stage against the CPU oracle."""
    sd = O.make_state_dict(manifest["wnet_em"], seed=7)
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    x = O.synthetic_diffractions(2, seed=17, H=H, W=W)
    ref = O.st_pipeline(sd, x)
    inf = S.SwinWNetInference(m, DEV)
    out = inf(x.to(DEV))
    assert out.shape == ref["images_masked_hr"].shape
    for k in ("seg_lr_logits", "upscaled_norm", "seg_hr_logits", "images_masked_hr"):
        assert relerr(getattr(inf, k), ref[k]) <= TOL, k


def test_empty_batch_returns_empty_result(wnet_em):
    inf = S.SwinWNetInference(wnet_em, DEV)
    out = inf(torch.zeros(0, 1, 250, 480, device=DEV))
    assert out.shape == (0, 2, 500, 960)


def test_batch_independence_and_determinism_full_size(wnet_em):
    """size-independent properties at the dataset geometry: a diffraction's result does not depend on its
    batch neighbours (how the batch is sharded over GPUs) and the path is run-to-run deterministic."""
    x = O.synthetic_diffractions(3, seed=21).to(DEV)
    inf = S.SwinWNetInference(wnet_em, DEV)
    full = inf(x).clone()
    again = inf(x).clone()
    assert torch.equal(full, again)
    solo = inf(x[1:2]).clone()
    assert torch.equal(full[1:2], solo)
    assert full.shape == (3, 2, 500, 960) and torch.isfinite(full).all()


def test_cuda_graph_replay_matches_eager_and_tracks_weights(manifest):
    """cuda_graph=True: replayed pipeline == eager pipeline bit for bit, for repeated calls, a second input shape, and
    after the parameters were replaced (the graph must be re-captured, not replayed with stale packed weights)."""
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(O.make_state_dict(manifest["wnet_em"], seed=1), strict=True)
    eager = S.SwinWNetInference(m, DEV)
    graphed = S.SwinWNetInference(m, DEV, cuda_graph=True)
    xa = O.synthetic_diffractions(2, seed=51, H=60, W=80, two_channel=False).to(DEV)
    xb = O.synthetic_diffractions(1, seed=52, H=40, W=100, two_channel=False).to(DEV)
    for x in (xa, xb, xa * 1.5, xa):
        ref = eager(x).clone()
        assert torch.equal(graphed(x), ref)
        assert torch.equal(graphed.seg_map_lr, eager.seg_map_lr)
    m.load_state_dict(O.make_state_dict(manifest["wnet_em"], seed=2), strict=True)
    ref = eager(xa).clone()
    assert torch.equal(graphed(xa), ref)


def test_run_host_pipelined_copies_match_device_call(wnet_em):
    """public host-data call (pinned host in -> pinned host out, chunked, copies overlapped with compute on side streams)
    returns exactly what the device call returns, chunk boundaries and ragged last chunk included."""
    x = O.synthetic_diffractions(5, seed=33, H=60, W=80, two_channel=False)
    inf = S.SwinWNetInference(wnet_em, DEV)
    ref = inf(x.to(DEV)).clone()
    xh = x.pin_memory()
    for chunk in (2, 5, 8):
        out = inf.run_host(xh, chunk=chunk)
        inf.host_done.synchronize()
        assert out.is_pinned() and out.shape == ref.shape
        assert torch.equal(out, ref.cpu())
    # full chunks replay a captured CUDA graph (host_graph, default): replays of the same graph on new data, the eager
    # host path, and the bound on the number of graphs kept
    assert 1 <= len(inf._graphs) <= inf.MAX_GRAPHS
    x2 = O.synthetic_diffractions(4, seed=34, H=60, W=80, two_channel=False)
    ref2 = inf(x2.to(DEV)).clone()
    out2 = inf.run_host(x2.pin_memory(), chunk=2)
    inf.host_done.synchronize()
    assert torch.equal(out2, ref2.cpu())
    assert torch.equal(inf.images_masked_hr, ref2[2:])            # cached stage attributes: those of the last chunk
    eager = S.SwinWNetInference(wnet_em, DEV, host_graph=False)
    out3 = eager.run_host(x2.pin_memory(), chunk=2)
    eager.host_done.synchronize()
    assert torch.equal(out3, out2) and not eager._graphs
    for chunk in (1, 3, 4):
        inf.run_host(x2.pin_memory(), chunk=chunk)
    inf.host_done.synchronize()
    assert len(inf._graphs) == inf.MAX_GRAPHS


def test_state_dict_round_trip_and_repack(manifest):
    sd = O.make_state_dict(manifest["unet"], seed=2)
    m = S.SwinUNet(depths=D2)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    x = O.normalize_piecewise(O.synthetic_diffractions(1, seed=9, H=20, W=30, two_channel=False))[0].to(DEV)
    with torch.no_grad():
        y1 = m(x).clone()
        sd2 = O.make_state_dict(manifest["unet"], seed=3)
        m.load_state_dict(sd2, strict=True)          # packed-weight caches must be rebuilt
        y2 = m(x).clone()
    assert relerr(y1, O.swin_unet(sd, x.cpu())) <= TOL
    assert relerr(y2, O.swin_unet(sd2, x.cpu())) <= TOL
    for k, v in m.state_dict().items():
        assert torch.equal(v.cpu(), sd2[k]), k


@pytest.fixture(scope="module")
def surrogate(manifest):
    """"trained-like" surrogate checkpoint (damped seeded body + fitted conv heads, oracle/make_surrogate_heads.py):
    segmentation probabilities are decisive (1.7 % of pixels within 0.05 of the threshold) and the SR output has
    diffraction peaks, so the mask / PSNR / physics gates are meaningful."""
    import os
    sd = O.surrogate_state_dict(manifest["wnet_em"], os.path.join(os.path.dirname(__file__), "golden", "surrogate_heads.pt"))
    m = S.SwinWNet(error_matrix=True, depths=D2)
    m.load_state_dict(sd, strict=True)
    return sd, m.to(DEV).eval()


def test_acceptance_gates_surrogate_checkpoint(surrogate):
    """north_star gates on the reference's evaluation protocols, CUDA path vs fp32 oracle, same inputs and weights:
    logits / images <= 2e-2 max-norm; binarised masks (>= 0.5, tests.py:12-16) identical on >= 99.9 % of pixels;
    PSNR (clamp(0,1), data_range 1, tests.py:332-357) within 0.05 dB; d-space integral- and peak-intensity distortion
    (tests.py:401-447 call pattern) within 1 % (batch mean)."""
    import numpy as np
    from oracle import diffraction_metrics_oracle as DM
    F = torch.nn.functional
    sd, model = surrogate
    B = 4
    images = O.synthetic_diffractions(B, seed=41)                     # [B,2,250,480]

    def protocol(seg1, upscale, seg2, dev):
        x = images.to(dev)
        seg, skips = seg1(x)
        xm = x * torch.sigmoid(seg)
        lr = F.interpolate(xm, scale_factor=0.5, mode="bilinear", align_corners=False)
        norm_lr, _ = O.normalize_piecewise(lr)
        norm_hr, params = O.normalize_piecewise(xm)
        sr, _ = upscale(norm_lr, skips)                               # physics / PSNR protocol: half-resolution input
        den = O.denormalize_piecewise(sr, params)
        up, skips_sr = upscale(norm_hr, skips)                        # inference pipeline: full resolution
        seg_hr, _ = seg2(O.denormalize_piecewise(up, params), skips_sr)
        return [t.float().cpu() for t in (seg, lr, sr, den, norm_hr, up, seg_hr)]

    with torch.no_grad():
        c = protocol(model.segment_1, model.upscale, model.segment_2, DEV)
        o = protocol(lambda x: O.segment_1(sd, x), lambda x, s: O.upscale(sd, x, s), lambda x, s: O.segment_2(sd, x, s), "cpu")
    names = ("seg_lr", "lr", "sr_half", "den", "norm_hr", "up_full", "seg_hr")
    for n, a, b in zip(names, c, o):
        e = relerr(a, b)
        print(f"max-norm rel err {n}: {e:.3e}")
        if n not in ("lr", "norm_hr"):        # glue intermediates (per-image min/max + log1p of image*sigmoid): reported only
            assert e <= TOL, n
    for n, a, b in (("LR", c[0], o[0]), ("HR", c[6], o[6])):
        agree = ((torch.sigmoid(a) >= 0.5) == (torch.sigmoid(b) >= 0.5)).float().mean().item()
        print(f"mask agreement {n}: {agree:.6f}  (foreground fraction {(torch.sigmoid(b) >= 0.5).float().mean().item():.3f})")
        assert agree >= (0.995 if BF16 else 0.999), (n, agree)
    p_c, p_o = psnr(c[2], o[4]), psnr(o[2], o[4])
    print("PSNR cuda / oracle:", p_c, p_o)
    assert abs(p_c - p_o) <= (0.2 if BF16 else 0.05)
    m_c = DM.physical_metrics(c[3].numpy(), o[1].numpy())
    m_o = DM.physical_metrics(o[3].numpy(), o[1].numpy())
    # Conditioning of the gauge itself: the metric goes through scipy.find_peaks thresholds and integer peak windows
    # (Diffraction_metrics.py:96-144), so on some samples it is bistable — the ORACLE's own value jumps by orders of
    # magnitude under a 1e-3 (max-norm) random perturbation of its own output, twenty times below the 2e-2 tolerance of
    # the images (measured: sample 3 of seed 41 flips between 2.264 and 161.36).  Such samples cannot gauge a 1 %
    # agreement; they are detected here with the oracle alone (never with the CUDA result) and left out of the mean.
    den_o = o[3]
    stable = np.ones(B, dtype=bool)
    for seed in range(8):
        g = torch.Generator().manual_seed(1000 + seed)
        noisy = den_o + torch.randn(den_o.shape, generator=g) * den_o.abs().max() * 1e-3
        m_n = DM.physical_metrics(noisy.numpy(), o[1].numpy())
        for k in ("Integral Intensity", "Peak Intensity"):
            stable &= np.abs(np.asarray(m_n[k]) - np.asarray(m_o[k])) <= 0.005 * np.abs(np.asarray(m_o[k])) + 1e-9
    print("well-conditioned samples for the physics gauge:", stable.tolist())
    assert stable.sum() >= 2, "physics gauge is ill-conditioned on almost every sample of this batch"
    for k in ("Integral Intensity", "Peak Intensity"):
        a, b = float(np.mean(np.asarray(m_c[k])[stable])), float(np.mean(np.asarray(m_o[k])[stable]))
        print(k, "cuda / oracle:", a, b, "per-sample", m_c[k], m_o[k])
        assert b > 0 and abs(a - b) <= (0.05 if BF16 else 0.01) * abs(b), (k, a, b)
