"""Pins oracle/diffraction_metrics_oracle.py: runs the UNMODIFIED reference Diffraction_metrics.py (authoring
container only) on seeded synthetic diffractions and stores its outputs in tests/golden/physics_metrics.json.
TEST INFRASTRUCTURE ONLY.   python oracle/make_golden_physics.py"""
import json
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
import Diffraction_metrics as RM  # noqa: E402  (reference)
from oracle import swinwnet_oracle as O  # noqa: E402
from oracle import diffraction_metrics_oracle as DM  # noqa: E402


def inputs(seed, B=3):
    hr = O.synthetic_diffractions(B, seed=seed, two_channel=False)
    g = torch.Generator().manual_seed(seed)
    pred = hr * (1.0 + 0.05 * torch.randn(hr.shape, generator=g)) * 1.03
    true = torch.nn.functional.interpolate(hr, scale_factor=0.5, mode="bilinear", align_corners=False)
    return pred, true


def main():
    calc = RM.DiffractionMetricsCalculator(DM.D_CENTERS_HR, DM.D_CENTERS_LR, device="cpu")
    out = {}
    for seed in (31, 32):
        pred, true = inputs(seed)
        m = calc(pred, true, peak_params_pred={"scale": True}, peak_params_true={"scale": False}, tol=0.05)
        di = calc.qw_pred.tensor_to_d(pred)
        out[str(seed)] = {"metrics": {k: [float(v) for v in vals] for k, vals in m.items()},
                          "I_sum": [float(s["I"].sum()) for s in di], "I_max": [float(s["I"].max()) for s in di],
                          "I_argmax": [int(s["I"].argmax()) for s in di]}
    with open(os.path.join(ROOT, "tests", "golden", "physics_metrics.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main()
