"""Pins the oracle (oracle/swinwnet_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/reference_outputs.npz, produced by oracle/make_golden.py in the authoring container)."""
import torch

from oracle import swinwnet_oracle as O

TOL = 2e-5  # fp32 restatement vs fp32 reference: max abs error relative to max(|ref|,1)


def close(a, b, tol=TOL):
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item() / max(1.0, b.abs().max().item())
    assert err <= tol, err


def test_case_A_multimodal_pipeline(manifest, golden):
    sd = O.make_state_dict(manifest["wnet_em"], seed=1)
    x = O.synthetic_diffractions(2, seed=3, H=40, W=60, two_channel=False)
    r = O.st_pipeline(sd, x)
    seg, skips = O.segment_1(sd, r["images"])
    close(seg, golden["A_seg_lr_logits"])
    for i, s in enumerate(skips):
        close(s, golden[f"A_skip{i}"])
    for k in ("seg_map_lr", "norm", "upscaled_norm", "upscaled_denorm", "seg_map_hr", "images_masked_hr"):
        close(r[k], golden["A_" + k])


def test_case_B_diffraction_only_manual(manifest, golden):
    sd = O.make_state_dict(manifest["wnet"], seed=1)
    x = O.synthetic_diffractions(1, seed=4, H=36, W=50, two_channel=False)
    r = O.st_pipeline(sd, x, two_channel=False)
    close(r["seg_lr_logits"], golden["B_seg_lr_logits"])
    close(r["upscaled_norm"], golden["B_upscaled_norm"])
    close(r["seg_hr_logits"], golden["B_seg_hr_logits"])


def test_case_B2_odd_image_patch_embed_pad(manifest, golden):
    sd = O.make_state_dict(manifest["wnet"], seed=1)
    x = O.synthetic_diffractions(1, seed=8, H=35, W=51, two_channel=False)
    seg, _ = O.segment_1(sd, x / 100.0)
    close(seg, golden["B2_seg_lr_logits"])


def test_case_C_single_branch(manifest, golden):
    x = O.synthetic_diffractions(1, seed=5, H=30, W=44, two_channel=False)
    xn, _ = O.normalize_piecewise(x)
    close(O.swin_unet(O.make_state_dict(manifest["unet"], seed=1), xn), golden["C_unet"])
    close(O.swin_unet_sr(O.make_state_dict(manifest["unetsr"], seed=1), xn), golden["C_unetsr"])


def test_case_D_even_step_geometry(manifest, golden):
    sd = O.make_state_dict(manifest["wnet_em"], seed=1)
    x = O.synthetic_diffractions(1, seed=6, H=40, W=60)
    xn, _ = O.normalize_piecewise(x)
    _, sk = O.segment_1(sd, xn)
    up, _ = O.upscale(sd, golden["D_lr_input"], sk)
    close(up, golden["D_upscaled"])


def test_case_E_full_geometry(manifest, golden):
    sd = O.make_state_dict(manifest["wnet_em"], seed=1)
    x = O.synthetic_diffractions(1, seed=7, two_channel=False)
    r = O.st_pipeline(sd, x)
    close(r["seg_lr_logits"][:, :, ::5, ::5], golden["E_seg_lr_logits_s5"])
    close(r["upscaled_norm"][:, :, ::5, ::5], golden["E_upscaled_norm_s5"])
    close(r["seg_map_hr"][:, :, ::5, ::5], golden["E_seg_map_hr_s5"])
    close(r["images_masked_hr"][:, :, ::5, ::5], golden["E_images_masked_hr_s5"], tol=2e-4)


def test_shift_mask_is_standard_swin():
    """shift>0 has no executable reference (SwinWNet.py:147 shape bug); check the oracle's
    mask semantics: tokens attend only within their region."""
    rid = O.shift_region_ids(10, 10, 5, 2)
    assert rid.unique().numel() == 9
    assert rid[0, 0] == 0 and rid[9, 9] == 8 and rid[5, 9] == 5


def test_physics_metric_oracle_vs_reference_golden():
    """oracle/diffraction_metrics_oracle.py against outputs of the unmodified reference Diffraction_metrics.py
    (tests/golden/physics_metrics.json, made by oracle/make_golden_physics.py)."""
    import json
    import os
    import numpy as np
    from oracle import diffraction_metrics_oracle as DM
    from oracle.make_golden_physics import inputs
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "physics_metrics.json")))
    for seed, g in gold.items():
        pred, true = inputs(int(seed))
        m = DM.physical_metrics(pred.numpy(), true.numpy())
        for k in ("Integral Intensity", "Peak Intensity", "Shape"):
            assert np.allclose(m[k], g["metrics"][k], rtol=2e-4, atol=1e-6), (seed, k, m[k], g["metrics"][k])
            assert sum(g["metrics"][k]) > 0          # the fixture exercises matched peaks
        for b in range(pred.shape[0]):
            _, I = DM.to_d_space(pred[b, 0].numpy(), DM.D_CENTERS_HR)
            assert abs(float(I.sum()) - g["I_sum"][b]) <= 1e-5 * g["I_sum"][b]
            assert int(I.argmax()) == g["I_argmax"][b]
