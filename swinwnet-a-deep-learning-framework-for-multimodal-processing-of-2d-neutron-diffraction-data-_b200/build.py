"""Builds libswinwnet_b200[_<variant>].so in-tree with nvcc for sm_100a (no torch involved).

Variants (selected at load time with SWN_LIB_VARIANT, see _lib.py):
    ""      the product build: IEEE fp16 tensor-core operands, fp32 accumulation
    "bf16"  -DSWN_OPERAND_BF16=1: bf16 operands (same kernels)
    "prof"  -DSWN_MLP_PROFILE=1 -DSWN_TUNING_HOOKS=1: in-kernel role-wait clocks + getenv tiling overrides (tools/ only)

The sha256 of the sources + flags is compiled INTO the library (swn_build_digest()); _lib.load() compares it with the
digest of the sources it sits next to, so a stale .so can never run silently against newer packing / Python code.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "rowgemm.cu", "rowgemm_persist.cu", "mlp.cu", "mlp_persist.cu", "window_attn.cu", "small_block.cu",
           "swin_fused.cu", "swin_warp.cu", "expand_warp.cu", "cross_attn.cu", "elementwise.cu", "train_ops.cu"]
HEADERS = ["common.cuh", "kernels.h", os.path.join("..", "..", "include", "swinwnet_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
VARIANTS = {"": [], "bf16": ["-DSWN_OPERAND_BF16=1"], "prof": ["-DSWN_MLP_PROFILE=1", "-DSWN_TUNING_HOOKS=1"],
            # scratch variants for A/B experiments on the GPU box (tools/gpu_call_ab.sh); flags come from SWN_X<n>_FLAGS at BUILD time
            "x1": os.environ.get("SWN_X1_FLAGS", "-DSWN_EXP=1").split(), "x2": os.environ.get("SWN_X2_FLAGS", "-DSWN_EXP=2").split(),
            "x3": os.environ.get("SWN_X3_FLAGS", "-DSWN_EXP=3").split()}


def lib_path(variant=""):
    return os.path.join(HERE, f"libswinwnet_b200{'_' + variant if variant else ''}.so")


def flags(variant=""):
    extra = [f for f in os.environ.get("SWN_NVCC_EXTRA", "").split() if f]   # ad-hoc experiment flags
    return NVCC_FLAGS + VARIANTS[variant] + extra


def digest(variant=""):
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(flags(variant)).encode())
    return h.hexdigest()


def build(force=False, verbose=False, variant=""):
    lib, stamp, dig = lib_path(variant), lib_path(variant) + ".stamp", digest(variant)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs, procs = [], []
    bdir = os.path.join(HERE, "build", variant or "default")
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags(variant), f'-DSWN_BUILD_DIGEST="{dig}"', "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-o", lib, *objs, "-lcudart"])
    with open(stamp, "w") as f:   # fast path only (git-ignored); the authoritative check is swn_build_digest()
        f.write(dig)
    return lib


if __name__ == "__main__":
    v = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    for variant in (v or [""]):
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant))
