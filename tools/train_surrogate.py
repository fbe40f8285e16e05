"""Surrogate "trained" multimodal checkpoint, produced by the REFERENCE'S OWN trainers (unmodified, imported from
baseline/_ref): SegmentatorTrainer -> UpscalerTrainer -> FullModelTrainer (Segmentator_pretrain.py, Upscaler_pretrain.py,
FullModel_supervised_trainer.py) for a few hundred optimizer steps on the six shipped diffractions + their masks
(datasets/*.npy, segmentation_maps.pkl rows 24/27/23/6/1/18, noise-augmented as in the notebooks: x + N(100, 20)) and
seeded synthetic Debye-Scherrer diffractions (benchdata.py) with thresholded peak masks.

Why: the shipped models/*.pth are git-LFS pointers, and with random weights sigmoid(seg) clusters on the 0.5 threshold,
so the mask-agreement / PSNR / physics gates of the north star only mean something on trained-like weights.  No
branch damping, no hand-fitted heads: whatever the reference's optimisation produces is the checkpoint.

Run on the GPU box (minutes):   python tools/train_surrogate.py --out gpurun_out/surrogate_wnet_em_fp16.pt
The state_dict is stored rounded to fp16-representable values (58 MB: gpurun returns at most 64 MiB); both sides of
every parity test load exactly these values as fp32 weights.  Keep it under tests/golden/_weights/ (git-ignored).
TEST INFRASTRUCTURE: nothing in the product package imports this.
"""
import argparse
import hashlib
import json
import os
import pickle
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import benchdata  # noqa: E402
from stage_reference import import_reference  # noqa: E402

REAL = [("Si", 24), ("UO2", 27), ("Rb", 23), ("C_graphite", 6), ("Al2O3_sapphire", 1), ("Na2Ca3Al2F14", 18)]


def real_samples(ref_dir):
    d = pickle.load(open(os.path.join(ref_dir, "datasets", "segmentation_maps.pkl"), "rb"))
    xs, ms = [], []
    for name, row in REAL:
        xs.append(torch.from_numpy(np.load(os.path.join(ref_dir, "datasets", f"{name}_diffraction.npy"))).float())
        ms.append(torch.from_numpy(np.asarray(d.iloc[row]["Mask"]).astype(np.int64)))
    return torch.stack(xs)[:, None], torch.stack(ms)


def build_dataset(ref_dir, n_syn, seed):
    """[N,1,250,480] fp32 images + [N,250,480] int64 masks: 6 real (x2 noise realisations) + n_syn synthetic."""
    g = torch.Generator().manual_seed(seed)
    xr, mr = real_samples(ref_dir)
    imgs, masks = [], []
    for rep in range(2):
        noise = (100.0 + 20.0 * torch.randn(xr.shape, generator=g)) if rep else torch.zeros_like(xr)
        imgs.append(xr + noise)
        masks.append(mr)
    xs = benchdata.synthetic_diffractions(n_syn, seed=seed + 7, two_channel=False)
    imgs.append(xs)
    masks.append((xs[:, 0] > 300.0).long())       # peaks stand clear of the N(100, 20) background
    return torch.cat(imgs), torch.cat(masks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "surrogate_wnet_em_fp16.pt"))
    ap.add_argument("--n-syn", type=int, default=36)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--epochs", type=int, nargs=3, default=[10, 5, 8], help="segmentator / upscaler / full-model epochs")
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--gamma0", type=float, default=0.3, help="cross-attention gamma before the joint stage (reference init is 0)")
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()

    ref_dir, R, _ = import_reference()
    from Segmentator_pretrain import SegmentatorTrainer
    from Upscaler_pretrain import UpscalerTrainer
    from FullModel_supervised_trainer import FullModelTrainer

    torch.manual_seed(a.seed)
    x, m = build_dataset(ref_dir, a.n_syn, a.seed)
    n_val = max(a.batch, len(x) // 8)
    perm = torch.randperm(len(x), generator=torch.Generator().manual_seed(a.seed))
    tr, va = perm[n_val:], perm[:n_val]
    mk = lambda idx, sh: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x[idx], m[idx]), batch_size=a.batch,
                                                     shuffle=sh, drop_last=sh, generator=torch.Generator().manual_seed(a.seed))
    train_loader, val_loader = mk(tr, True), mk(va, False)
    model = R.SwinWNet(error_matrix=True, depths=[2, 2, 2, 2]).to(a.device)
    log = {"device": a.device, "n_train": len(tr), "n_val": len(va), "batch": a.batch, "epochs": a.epochs, "lr": a.lr,
           "torch": torch.__version__, "stages": {}}
    t0 = time.time()
    fp16 = a.device.startswith("cuda")
    e1, e2, e3 = a.epochs
    if e1:
        t = SegmentatorTrainer(model, train_loader, val_loader, a.device, num_epochs=e1, warmup_epochs=min(2, e1 - 1), lr=a.lr, use_fp16=fp16, verbose=True)
        log["stages"]["segmentator"] = t.train()
        t.release_training_state()
    if e2:
        t = UpscalerTrainer(model, train_loader, val_loader, a.device, num_epochs=e2, warmup_epochs=min(1, e2 - 1), lr=a.lr, use_fp16=fp16, verbose=True)
        log["stages"]["upscaler"] = t.train()
        t.release_training_state()
    for p in model.parameters():
        p.requires_grad = True
    with torch.no_grad():
        for blk in list(model.ca_seg_to_sr.blocks) + list(model.ca_sr_to_seg.blocks):
            blk.gamma.fill_(a.gamma0)
    if e3:
        t = FullModelTrainer(model, train_loader, val_loader, a.device, num_epochs=e3, warmup_epochs=min(2, e3 - 1), lr=a.lr / 3, verbose=True)
        t.train()
        t.release_training_state()
    log["train_seconds"] = time.time() - t0
    sd = {k: (v.detach().cpu().half() if v.is_floating_point() else v.detach().cpu()) for k, v in model.state_dict().items()}
    bad = [k for k, v in sd.items() if v.is_floating_point() and not torch.isfinite(v).all()]
    if bad:
        raise SystemExit(f"non-finite parameters after training: {bad[:5]}")
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    torch.save(sd, a.out)
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    log["state_dict_sha256"] = h.hexdigest()
    log["gammas"] = {k: float(v) for k, v in sd.items() if k.endswith("gamma")}
    # a first look at what the gates will see: how far from the 0.5 threshold do the LR probabilities sit?
    model.eval()
    with torch.no_grad():
        seg, _ = model.segment_1(torch.cat([x[va], torch.sqrt(torch.abs(x[va]))], 1).to(a.device))
        pr = torch.sigmoid(seg.float())
        log["val_frac_within_0.02_of_threshold"] = float(((pr - 0.5).abs() < 0.02).float().mean())
        log["val_mask_fraction"] = float((pr > 0.5).float().mean())
        log["val_pixel_acc"] = float(((pr[:, 0] > 0.5).cpu() == (m[va] > 0)).float().mean())
    with open(os.path.splitext(a.out)[0] + ".json", "w") as f:
        json.dump(log, f, indent=1)
    print(json.dumps({k: v for k, v in log.items() if k != "stages"}))


if __name__ == "__main__":
    main()
