"""Checkpoint ingestion for the drop-in modules: the tolerant loader of the reference's viewer
(inference_gui/swinwnet_viewer_gui.py:129-151 — nested ``state_dict`` / ``model_state_dict``, DataParallel ``module.``
prefixes, modality inferred from the patch-embed weight) plus depth inference from the block keys, so that checkpoints of
either depth configuration load with ``strict=True`` instead of the viewer's ``strict=False``."""
import os
import re

import numpy as np
import torch


def load_state_dict_any(pth_path_or_obj, map_location="cpu") -> dict:
    ckpt = torch.load(pth_path_or_obj, map_location=map_location) if isinstance(pth_path_or_obj, (str, bytes, os.PathLike)) or hasattr(
        pth_path_or_obj, "read") else pth_path_or_obj
    if not isinstance(ckpt, dict):
        raise ValueError("Unsupported checkpoint format")
    if isinstance(ckpt.get("state_dict"), dict):
        sd = ckpt["state_dict"]
    elif isinstance(ckpt.get("model_state_dict"), dict):
        sd = ckpt["model_state_dict"]
    else:
        sd = ckpt
    if any(k.startswith("module.") for k in sd.keys()):
        sd = {k.replace("module.", "", 1): v for k, v in sd.items()}
    return sd


def infer_error_matrix_flag_from_sd(sd: dict) -> bool:
    k = "patch_embed.proj.weight"
    return bool(k in sd and hasattr(sd[k], "shape") and int(sd[k].shape[1]) >= 2)


def infer_depths_from_sd(sd: dict, encoder_prefix=None):
    """depths list of the encoder (= constructor argument) from the ``<encoder>.layers.<i>.blocks.<j>.`` keys."""
    if encoder_prefix is None:
        encoder_prefix = "segmentator_encoder" if any(k.startswith("segmentator_encoder.") for k in sd) else "encoder"
    pat = re.compile(r"^" + re.escape(encoder_prefix) + r"\.layers\.(\d+)\.blocks\.(\d+)\.")
    depth = {}
    for k in sd:
        m = pat.match(k)
        if m:
            i, j = int(m.group(1)), int(m.group(2))
            depth[i] = max(depth.get(i, 0), j + 1)
    if not depth:
        raise ValueError(f"no '{encoder_prefix}.layers.*.blocks.*' keys in the checkpoint")
    return [depth[i] for i in range(max(depth) + 1)]


def build_model_from_checkpoint(pth_path_or_obj, device=None):
    """SwinWNet / SwinUNet / SwinUNetSR (by key prefixes) with inferred depths and modality, loaded strict=True."""
    from . import model as M
    sd = load_state_dict_any(pth_path_or_obj)
    depths = infer_depths_from_sd(sd)
    if any(k.startswith("segmentator_encoder.") for k in sd):
        net = M.SwinWNet(error_matrix=infer_error_matrix_flag_from_sd(sd), depths=depths)
    elif any(k.startswith("head.reconstruction.") for k in sd):
        net = M.SwinUNetSR(depths=depths)
    else:
        net = M.SwinUNet(depths=depths)
    net.load_state_dict(sd, strict=True)
    return net.to(device) if device is not None else net


# ---------------------------------------------------------------------------------------------------------------------
# .npy batch ingestion (the second half of §8 f-4; inference_gui/swinwnet_viewer_gui.py:113-123 `_as_4d`, :598-606
# `load_npy`): a file holds either one array or a pickled dict of named arrays; 2-D [H,W], 3-D [B,H,W] and 4-D [B,C,H,W]
# arrays are coerced to the [B,C,H,W] layout the pipeline takes.
# ---------------------------------------------------------------------------------------------------------------------
def as_4d(x) -> np.ndarray:
    x = np.asarray(x)
    if x.ndim == 2:
        return x[None, None, ...]
    if x.ndim == 3:
        return x[:, None, ...]
    if x.ndim == 4:
        return x
    raise ValueError(f"Unsupported array shape: {x.shape}")


def load_npy_batch(paths, key="images", pin=True) -> torch.Tensor:
    """One or several ``.npy`` files -> one fp32 host tensor [B,C,H,W] (pinned by default, ready for
    ``SwinWNetInference.run_host``).  Dict files (``np.save`` of a dict, as the viewer writes) contribute ``key``; all
    files must agree on (C,H,W).  Raises ``ValueError`` on shape mismatches, like the viewer's loader."""
    if isinstance(paths, (str, bytes, os.PathLike)):
        paths = [paths]
    parts = []
    for p in paths:
        obj = np.load(p, allow_pickle=True)
        item = obj.item() if getattr(obj, "shape", None) == () else obj
        if isinstance(item, dict):
            if key not in item:
                raise ValueError(f"{p}: no '{key}' entry (has {sorted(item)})")
            item = item[key]
        a = as_4d(item).astype(np.float32, copy=False)
        if parts and a.shape[1:] != parts[0].shape[1:]:
            raise ValueError(f"{p}: shape {a.shape[1:]} does not match {parts[0].shape[1:]}")
        parts.append(a)
    if not parts:
        raise ValueError("no input files")
    t = torch.from_numpy(np.ascontiguousarray(np.concatenate(parts, 0)))
    return t.pin_memory() if pin and torch.cuda.is_available() else t
