"""Surrogate "trained-like" weights for the acceptance gates that need a meaningful model output
(mask agreement @0.5, PSNR, d-space physics metrics): the shipped checkpoints are git-LFS pointers in the
reference mount, and with purely random weights the segmentation probabilities cluster at 0.5 and
the SR output has no diffraction peaks.

Recipe (deterministic, CPU, minutes): keep the seeded random body of
``make_state_dict(manifest['wnet_em'], 1, branch_scale=0.1)`` (attention / MLP branch outputs damped so the residual
stream carries the input through the network) and fit ONLY the two conv heads (~12k parameters) on features of the frozen body, with the reference's own
training objectives: BCE on a peak mask for ``segmentator_head`` (Segmentator_pretrain.py) and MSE between
``upscale(normalised half-resolution input)`` and the normalised full-resolution image for
``upscaler_head.reconstruction`` (Upscaler_pretrain.py / tests.py:332-357).  The 8 fitted tensors are stored in
``tests/golden/surrogate_heads.pt`` (~60 KB).  TEST INFRASTRUCTURE ONLY (uses the oracle).

    python oracle/make_surrogate_heads.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import swinwnet_oracle as O  # noqa: E402

F = torch.nn.functional
SEG_KEYS = ["segmentator_head.seg_head.0.weight", "segmentator_head.seg_head.0.bias",
            "segmentator_head.seg_head.2.weight", "segmentator_head.seg_head.2.bias"]
REC_KEYS = ["upscaler_head.reconstruction.0.weight", "upscaler_head.reconstruction.0.bias",
            "upscaler_head.reconstruction.2.weight", "upscaler_head.reconstruction.2.bias"]


def peak_mask(x):
    """pixels clearly above the N(100,20) background of the synthetic generator"""
    return (x[:, :1] > 300.0).float()


def seg_trunk(sd, x):
    t, pres = O.patch_embed(sd, "patch_embed.", x, 1)
    res = (pres[0] // 2, pres[1] // 2)
    skips, rl, bres = O.encoder(sd, "segmentator_encoder.", t, res, O.DEPTHS, O.HEADS)
    xb = O.bottleneck(sd, "segmentator_bottleneck.", skips[-1], bres, O.HEADS[-1])
    xd, _ = O.decoder(sd, "segmentator_decoder.", xb, bres, skips, rl, O.DEPTHS, O.HEADS)
    return xd, pres, skips


def sr_trunk(sd, x, skips_seg):
    t, pres = O.patch_embed(sd, "patch_embed.", x, 1)
    res = (pres[0] // 2, pres[1] // 2)
    skips, rl, bres = O.encoder(sd, "upscaler_encoder.", t, res, O.DEPTHS, O.HEADS)
    skips[-2], skips[-1] = O.multi_scale_cross_attention(sd, "ca_seg_to_sr.", [skips[-2], skips[-1]],
                                                         [skips_seg[-2], skips_seg[-1]])
    xb = O.bottleneck(sd, "upscaler_bottleneck.", skips[-1], bres, O.HEADS[-1])
    xd, _ = O.decoder(sd, "upscaler_decoder.", xb, bres, skips, rl, O.DEPTHS, O.HEADS)
    r = (pres[0] // 2, pres[1] // 2)
    for i in range(2):
        xd, r = O.patch_expanding(sd, f"upscaler_head.ups.{i}.", xd, r)
        xd = O.basic_layer(sd, f"upscaler_head.swin_blocks.{i}.", xd, r, 2, 3)
    return xd.view(x.shape[0], r[0], r[1], -1), r


def fit(params, loss_fn, steps, lr):
    opt = torch.optim.Adam(params, lr=lr)
    for it in range(steps):
        opt.zero_grad()
        loss = loss_fn()
        loss.backward()
        opt.step()
        if it % 50 == 0 or it == steps - 1:
            print(f"  step {it:4d} loss {loss.item():.5f}", flush=True)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    sd = O.make_state_dict(man["wnet_em"], seed=1, branch_scale=0.1)
    x = O.synthetic_diffractions(6, seed=50)                                  # [6,2,250,480]
    with torch.no_grad():
        xd, pres, skips = seg_trunk(sd, x)
    # ---- stage 1: segmentation head (BCE on the peak mask) ----
    print("fitting segmentator_head")
    seg_p = {k: sd[k].clone().requires_grad_(True) for k in SEG_KEYS}
    target = peak_mask(x)

    def seg_loss():
        logits = O.segmentation_head({**sd, **seg_p}, "segmentator_head.", xd, pres, 1)
        return F.binary_cross_entropy_with_logits(logits, target)
    fit(list(seg_p.values()), seg_loss, 300, 3e-3)
    sd.update({k: v.detach() for k, v in seg_p.items()})
    # ---- stage 2: reconstruction head on the reference's upscaler objective ----
    print("fitting upscaler_head.reconstruction")
    with torch.no_grad():
        seg = O.segmentation_head(sd, "segmentator_head.", xd, pres, 1)
        xm = x * torch.sigmoid(seg)
        lr_img = F.interpolate(xm, scale_factor=0.5, mode="bilinear", align_corners=False)
        norm_lr, _ = O.normalize_piecewise(lr_img)
        norm_hr, _ = O.normalize_piecewise(xm)
        feat, r = sr_trunk(sd, norm_lr, skips)                               # [6,252,480,12]
    rec_p = {k: sd[k].clone().requires_grad_(True) for k in REC_KEYS}

    def rec_loss():
        h = O.gelu_erf(O.conv3x3_nhwc(feat, rec_p[REC_KEYS[0]], rec_p[REC_KEYS[1]]))
        w2 = rec_p[REC_KEYS[2]]
        out = (h @ w2.view(w2.shape[0], -1).t() + rec_p[REC_KEYS[3]]).permute(0, 3, 1, 2)
        return F.mse_loss(out[:, :, :norm_hr.shape[2], :norm_hr.shape[3]], norm_hr) * 100.0
    fit(list(rec_p.values()), rec_loss, 300, 3e-3)
    sd.update({k: v.detach() for k, v in rec_p.items()})
    # ---- stage 3: refit the (shared) segmentation head on LR + HR features, as FullModelTrainer's odd steps do ----
    print("refitting segmentator_head on LR + HR features")
    with torch.no_grad():
        norm_hr2, params = O.normalize_piecewise(xm)
        up, skips_sr = O.upscale(sd, norm_hr2, skips)
        den = O.denormalize_piecewise(up, params)
        t, pres2 = O.patch_embed(sd, "patch_embed.", den, 2)
        res2 = (pres2[0] // 4, pres2[1] // 4)
        sk2, rl2, bres2 = O.encoder(sd, "segmentator_encoder.", t, res2, O.DEPTHS, O.HEADS)
        sk2[-2], sk2[-1] = O.multi_scale_cross_attention(sd, "ca_sr_to_seg.", [sk2[-2], sk2[-1]], [skips_sr[-2], skips_sr[-1]])
        xb2 = O.bottleneck(sd, "segmentator_bottleneck.", sk2[-1], bres2, O.HEADS[-1])
        xd_hr, _ = O.decoder(sd, "segmentator_decoder.", xb2, bres2, sk2, rl2, O.DEPTHS, O.HEADS)
        target_hr = peak_mask(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))
    seg_p = {k: sd[k].clone().requires_grad_(True) for k in SEG_KEYS}

    def seg_loss2():
        w = {**sd, **seg_p}
        lo = O.segmentation_head(w, "segmentator_head.", xd, pres, 1)
        hi = O.segmentation_head(w, "segmentator_head.", xd_hr, pres2, 2)
        return F.binary_cross_entropy_with_logits(lo, target) + F.binary_cross_entropy_with_logits(hi, target_hr)
    fit(list(seg_p.values()), seg_loss2, 200, 3e-3)
    sd.update({k: v.detach() for k, v in seg_p.items()})
    out = {k: sd[k].clone() for k in SEG_KEYS + REC_KEYS}
    torch.save(out, os.path.join(ROOT, "tests", "golden", "surrogate_heads.pt"))
    print("saved", sum(v.numel() for v in out.values()), "parameters")


if __name__ == "__main__":
    main()
