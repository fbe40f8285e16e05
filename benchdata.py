"""Deterministic synthetic inputs and weights shared by bench.py, the tests and the oracle.

Neutral data-generation helpers (no model math): seeded Debye-Scherrer-like diffractions on the dataset grid
 and construction-order-independent weights for a key->shape manifest of the reference modules
(tests/golden/manifest.json).  There is no network for datasets or checkpoints, so both bench arms use these."""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]
WS = 5


def rel_pos_index(ws: int = WS) -> Tensor:
    """index[i,j] = (yi-yj+ws-1)*(2ws-1) + (xi-xj+ws-1)  (SwinWNet.py:163-172)."""
    t = torch.arange(ws * ws)
    y, x = t // ws, t % ws
    return (y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)


def ensure_2ch(x: Tensor) -> Tensor:
    return x if x.size(1) == 2 else torch.cat([x, torch.sqrt(torch.abs(x))], dim=1)


def synthetic_diffractions(B: int, seed: int = 0, H: int = 250, W: int = 480, two_channel: bool = True) -> Tensor:
    """Seeded Debye-Scherrer-like synthetic inputs on the dataset grid."""
    lam = torch.linspace(0.1, 10.0, H)
    th = torch.deg2rad(torch.linspace(-170.0, 170.0, W))
    d = lam[:, None] / (2.0 * torch.sin(torch.abs(th)[None, :] * 0.5))
    out = torch.empty(B, 1, H, W)
    for b in range(B):
        g = torch.Generator().manual_seed(seed * 100003 + b)
        K = int(torch.randint(8, 31, (1,), generator=g))
        dk = 0.5 + 6.5 * torch.rand(K, generator=g)
        amp = torch.exp(math.log(2e2) + (math.log(1.5e4) - math.log(2e2)) * torch.rand(K, generator=g))
        wk = 0.005 + 0.015 * torch.rand(K, generator=g)
        img = torch.zeros(H, W)
        for k in range(K):
            img += amp[k] * torch.exp(-0.5 * ((d - dk[k]) / (wk[k] * dk[k])) ** 2)
        img += (100.0 + 20.0 * torch.randn(H, W, generator=g)).clamp_min(1.0)
        out[b, 0] = img
    return ensure_2ch(out) if two_channel else out


def make_state_dict(manifest: Dict[str, Sequence[int]], seed: int = 0, branch_scale: float = 1.0) -> SD:
    """Deterministic, construction-order-independent weights for a key->shape manifest
    (tests/golden/manifest.json).  Magnitudes are "trained-like": LayerNorm affine is
    perturbed, the relative-position tables and every cross-attention ``gamma`` are
    non-trivial (the reference initialises gamma to 0, SwinWNet.py:776, which would
    hide cross-attention bugs).  ``branch_scale`` < 1 damps the output projections of every attention / MLP
    branch (a residual-dominant body, used by the surrogate checkpoint of oracle/make_surrogate_heads.py)."""
    sd: SD = {}
    for i, key in enumerate(sorted(manifest)):
        shape = tuple(manifest[key])
        g = torch.Generator().manual_seed(seed * 1000003 + i)
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("relative_position_index"):
            sd[key] = rel_pos_index(WS).clone()
        elif leaf == "gamma":
            sd[key] = torch.full(shape, 0.5)
        elif key.endswith("relative_position_bias_table"):
            sd[key] = 0.5 * torch.randn(shape, generator=g)
        elif re_norm(key) and leaf == "weight":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif re_norm(key) and leaf == "bias":
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif leaf in ("weight", "in_proj_weight"):
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        else:  # biases
            sd[key] = 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        if branch_scale != 1.0 and key.endswith(("attn.proj.weight", "attn.proj.bias", "mlp.3.weight", "mlp.3.bias",
                                                 "attn.out_proj.weight", "attn.out_proj.bias")):
            sd[key] = sd[key] * branch_scale
    return sd


def surrogate_state_dict(manifest: Dict[str, Sequence[int]], heads_path: str) -> SD:
    """the surrogate "trained-like" multimodal checkpoint: damped seeded body + fitted conv heads
    (tests/golden/surrogate_heads.pt, see oracle/make_surrogate_heads.py)."""
    sd = make_state_dict(manifest, seed=1, branch_scale=0.1)
    sd.update(torch.load(heads_path))
    return sd


def re_norm(key: str) -> bool:
    parts = key.split(".")
    return len(parts) >= 2 and parts[-2].startswith("norm")
