"""ORACLE (test infrastructure only) for the physics acceptance gauge of the SwinWNet hot path: the d-space
integral-intensity / peak-intensity / peak-shape distortion metrics of the reference's
``Diffraction_metrics.py`` (north_star: "within 1 % of the reference").

numpy + scipy restatement; every function cites the reference lines it follows.  Pinned against the reference
itself by ``oracle/make_golden.py`` -> ``tests/golden/physics_metrics.json`` -> ``tests/test_oracle_golden.py``.
Never imported by the product package.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
from scipy.signal import find_peaks

# evaluation grids of the reference protocol (tests.py:168-172)
D_CENTERS_HR = np.linspace(0.05318052, 7.49710258, 1241)
D_CENTERS_LR = np.linspace(0.0546658, 7.49180085, 832)


def to_d_space(img2d: np.ndarray, centers: Sequence[float], theta_range=(-170.0, 170.0), l_range=(0.1, 10.0)):
    """2-D (lambda x theta) intensity -> 1-D I(d) histogram (Qwrapper.tensor_to_d, Diffraction_metrics.py:35-70).

    d = L / (2 sin(|theta|/2)); pixels with d <= 7.5 are summed into the bin whose (mid-point) edges bracket d."""
    import torch  # fp32 torch ops so that pixels on bin edges fall exactly where the reference puts them
    c = torch.tensor(np.asarray(centers), dtype=torch.float32)
    H, W = img2d.shape
    edges = torch.zeros(len(c) + 1, dtype=torch.float32)
    edges[1:-1] = (c[:-1] + c[1:]) * 0.5
    edges[0] = c[0] - (c[1] - c[0]) * 0.5
    edges[-1] = c[-1] + (c[-1] - c[-2]) * 0.5
    theta = torch.deg2rad(torch.linspace(theta_range[0], theta_range[1], W))
    lam = torch.linspace(l_range[0], l_range[1], H)
    d = lam[:, None] / (2 * torch.sin(torch.abs(theta)[None, :] * 0.5))
    mask = d <= 7.5
    idx = (torch.bucketize(d[mask], edges) - 1).clamp(0, len(c) - 1)
    I = torch.zeros(len(c), dtype=torch.float32)
    I.scatter_add_(0, idx, torch.as_tensor(np.ascontiguousarray(img2d), dtype=torch.float32)[mask])
    return c.numpy(), I.numpy()


def peaks_of(d: np.ndarray, I: np.ndarray, scale: bool = False, height=0.05, distance=10, prominence=0.1, width=5,
             scale_factor=1.5, default_window=15) -> List[dict]:
    """find_peaks_for_batch for one sample (Diffraction_metrics.py:96-144, window rule :75-92)."""
    if scale:
        I = I / 4
    peaks, props = find_peaks(I, height=height, distance=distance, prominence=prominence, width=width)
    out = []
    for n, pk in enumerate(peaks):
        window = int(props["widths"][n] * scale_factor) if "widths" in props else default_window
        lo, hi = max(pk - window, 0), min(pk + window, len(d))
        dw, Iw = d[lo:hi], I[lo:hi]
        out.append({"d": float(d[pk]), "d_com": float(np.sum(dw * Iw) / np.sum(Iw)),
                    "integral_intensity": float(np.sum(Iw)), "max_intensity": float(I[pk]),
                    "profile_d": dw, "profile_I": Iw})
    return out


def _resample(d, I, d_center, x_ref):
    s = np.sum(I)
    if s <= 0:
        return None
    return np.interp(x_ref, (d - d_center) / d_center, I / s, left=0.0, right=0.0)


def _emd_shape(p1, p2, x_ref, eps=1e-12):
    """emd_shape_loss (Diffraction_metrics.py:150-203)."""
    a = _resample(p1["profile_d"], p1["profile_I"], p1["d"], x_ref)
    b = _resample(p2["profile_d"], p2["profile_I"], p2["d"], x_ref)
    if a is None or b is None:
        return 0.0
    a, b = np.maximum(a, 0), np.maximum(b, 0)
    a = a / (np.sum(a) + eps)
    b = b / (np.sum(b) + eps)
    return float(np.sum(np.abs(np.cumsum(a) - np.cumsum(b))) * (x_ref[1] - x_ref[0]))


def compare_peaks(pred: List[dict], true: List[dict], tol: float = 0.05):
    """compare_peak_sets (Diffraction_metrics.py:209-251): nearest true peak by d, matched if |d_com diff| <= tol;
    sums of squared log-ratios of integral / max intensity and EMD of the normalised profiles."""
    tot_i = tot_m = tot_s = 0.0
    if not pred or not true:
        return tot_i, tot_m, tot_s
    x_ref = np.linspace(-0.03, 0.03, 64)
    for p1 in pred:
        p2 = min(true, key=lambda q: abs(q["d"] - p1["d_com"]))
        if abs(p1["d_com"] - p2["d_com"]) > tol:
            continue
        tot_i += (math.log(max(p1["integral_intensity"], 0) + 1) - math.log(max(p2["integral_intensity"], 0) + 1)) ** 2
        tot_m += (math.log(max(p1["max_intensity"], 0) + 1) - math.log(max(p2["max_intensity"], 0) + 1)) ** 2
        tot_s += _emd_shape(p1, p2, x_ref)
    return tot_i, tot_m, tot_s


def physical_metrics(pred_2d: np.ndarray, true_2d: np.ndarray, centers_pred=D_CENTERS_HR, centers_true=D_CENTERS_LR,
                     tol: float = 0.05) -> Dict[str, List[float]]:
    """DiffractionMetricsCalculator.__call__ with the tests.py:441-447 call pattern (pred peaks scaled by 1/4).
    pred_2d / true_2d: [B, C, H, W] arrays; channel 0 is used (Diffraction_metrics.py:58)."""
    out = {"Integral Intensity": [], "Peak Intensity": [], "Shape": []}
    for b in range(pred_2d.shape[0]):
        dp, Ip = to_d_space(np.asarray(pred_2d[b, 0], dtype=np.float32), centers_pred)
        dt, It = to_d_space(np.asarray(true_2d[b, 0], dtype=np.float32), centers_true)
        i, m, s = compare_peaks(peaks_of(dp, Ip, scale=True), peaks_of(dt, It, scale=False), tol)
        out["Integral Intensity"].append(i)
        out["Peak Intensity"].append(m)
        out["Shape"].append(s)
    return out
