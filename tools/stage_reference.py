"""Stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

The reference is a flat directory of scripts (no setup.py / pyproject.toml), so "installing" it is a verbatim copy of
the files the hot path, the trainers and the physics gauge need, plus the six shipped diffractions and their masks.
Nothing under baseline/_ref/ is product source and nothing in the product package imports it; it is used by
  * bench.py --impl reference  (the reference's own CPU forward, kind "reference"),
  * tools/library_bar.py       (the reference in eager fp32 / TF32 / bf16-autocast on the B200: the library bar),
  * tools/train_surrogate.py   (the reference's own trainers produce the surrogate checkpoint of the parity gates).

    python tools/stage_reference.py [--src /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
PY_FILES = ["SwinWNet.py", "ST_Inference_Pipline.py", "Diffraction_metrics.py", "supervised_losses.py",
            "Segmentator_pretrain.py", "Upscaler_pretrain.py", "FullModel_supervised_trainer.py", "LICENSE"]
DATA_FILES = ["datasets/Si_diffraction.npy", "datasets/UO2_diffraction.npy", "datasets/Rb_diffraction.npy",
              "datasets/C_graphite_diffraction.npy", "datasets/Al2O3_sapphire_diffraction.npy",
              "datasets/Na2Ca3Al2F14_diffraction.npy", "datasets/segmentation_maps.pkl"]


def stage(src="/root/reference", quiet=False):
    """copy the files (only when the source tree exists); returns the manifest {relative path: sha256}."""
    if not os.path.isdir(src):
        return None
    man = {}
    for rel in PY_FILES + DATA_FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        data = open(s, "rb").read()
        man[rel] = hashlib.sha256(data).hexdigest()
        if not (os.path.exists(d) and open(d, "rb").read() == data):
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": man}, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"staged {len(man)} reference files under {DST}")
    return man


def ref_dir():
    """where the unmodified reference can be imported from at run time (None if nowhere)."""
    for p in (DST, "/root/reference"):
        if os.path.exists(os.path.join(p, "SwinWNet.py")):
            return p
    return None


def import_reference():
    """import the reference modules (SwinWNet, ST_Inference_Pipline, ...) from baseline/_ref or /root/reference."""
    p = ref_dir()
    if p is None:
        raise RuntimeError("the reference is staged neither under baseline/_ref nor under /root/reference")
    sys.dont_write_bytecode = True
    if p not in sys.path:
        sys.path.insert(0, p)
    import importlib
    return p, importlib.import_module("SwinWNet"), importlib.import_module("ST_Inference_Pipline")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    if stage(a.src) is None:
        raise SystemExit(f"{a.src} does not exist")
