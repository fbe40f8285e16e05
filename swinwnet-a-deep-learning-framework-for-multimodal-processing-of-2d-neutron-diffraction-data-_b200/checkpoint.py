"""Checkpoint ingestion for the drop-in modules: the tolerant loader of the reference's viewer
(inference_gui/swinwnet_viewer_gui.py:129-151 — nested ``state_dict`` / ``model_state_dict``, DataParallel ``module.``
prefixes, modality inferred from the patch-embed weight) plus depth inference from the block keys, so that checkpoints of
either depth configuration load with ``strict=True`` instead of the viewer's ``strict=False``."""
import re

import torch


def load_state_dict_any(pth_path_or_obj, map_location="cpu") -> dict:
    ckpt = torch.load(pth_path_or_obj, map_location=map_location) if isinstance(pth_path_or_obj, (str, bytes)) or hasattr(
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
