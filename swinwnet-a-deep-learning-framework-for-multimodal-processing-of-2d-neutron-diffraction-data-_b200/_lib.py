"""ctypes binding of libswinwnet_b200.so (C ABI: include/swinwnet_b200.h).

No fallback of any kind: if the library is missing or a call fails, a RuntimeError is raised."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SWN_LIB_VARIANT selects an alternative in-tree build of the same sources (build.py: "bf16" = bf16 tensor-core
# operands, "prof" = profiling / tuning build used by tools/); the default is the product build (fp16 operands).
VARIANT = os.environ.get("SWN_LIB_VARIANT", "")
LIB_PATH = os.path.join(HERE, f"libswinwnet_b200{'_' + VARIANT if VARIANT else ''}.so")

c_int, c_float, c_void_p, c_longlong, c_double = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_double


class RowGemmArgs(ctypes.Structure):
    """mirror of swn_rowgemm_args"""
    _fields_ = [
        ("A", c_void_p), ("a_mode", c_int), ("M", c_int), ("K", c_int), ("lda", c_int),
        ("ln_w", c_void_p), ("ln_b", c_void_p), ("ln_eps", c_float),
        ("gH", c_int), ("gW", c_int), ("gC", c_int), ("gHo", c_int), ("gWo", c_int),
        ("Wp", c_void_p), ("NT", c_int), ("nchunks", c_int), ("n_valid", c_int),
        ("e_mode", c_int), ("bias", c_void_p), ("out", c_void_p), ("ldo", c_int),
        ("res", c_void_p), ("ldres", c_int), ("alpha", c_void_p),
        ("xH", c_int), ("xW", c_int), ("xHs", c_int), ("xWs", c_int), ("ln2_w", c_void_p), ("ln2_b", c_void_p),
    ]


# every symbol include/swinwnet_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "swn_last_error": (ctypes.c_char_p, []),
    "swn_abi_version": (c_int, []),
    "swn_build_digest": (ctypes.c_char_p, []),
    "swn_sizeof_rowgemm_args": (c_int, []),
    "swn_operand_is_bf16": (c_int, []),
    "swn_mlp_config": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "swn_rowgemm": (c_int, [ctypes.POINTER(RowGemmArgs), c_void_p]),
    "swn_mlp": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                        c_void_p]),
    "swn_swin_block_small": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                     ctypes.POINTER(c_void_p), c_void_p]),
    "swn_swin_block_fused": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                     c_int, c_void_p]),
    "swn_swin_block_warp": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                    c_int, c_void_p]),
    "swn_set_phase_profile": (c_int, [c_void_p]),
    "swn_window_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p]),
    "swn_window_attention_frags": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                           c_void_p]),
    "swn_cross_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "swn_patch_embed": (c_int, [c_void_p] * 6 + [c_int] * 7 + [c_void_p]),
    "swn_seg_head": (c_int, [c_void_p] * 7 + [c_int] * 6 + [c_void_p]),
    "swn_recon_head": (c_int, [c_void_p] * 6 + [c_int] * 6 + [c_void_p]),
    "swn_copy_cols": (c_int, [c_void_p, c_int, c_void_p, c_int, c_longlong, c_int, c_void_p]),
    "swn_sigmoid_mask": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_void_p]),
    "swn_dspace_histogram": (c_int, [c_void_p, c_longlong, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "swn_normalize": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p]),
    "swn_ensure_2ch": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "swn_adamw_multi": (c_int, [c_void_p, c_void_p, c_int, c_double, c_double, c_double, c_double, c_double, c_double, c_void_p]),
    "swn_grad_bucket_copy": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_double, c_void_p]),
}

_lib = None


def load():
    """Load the shared library (built in-tree by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the SwinWNet forward.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
            fn.restype, fn.argtypes = res, args
        if lib.swn_abi_version() != 1 or lib.swn_sizeof_rowgemm_args() != ctypes.sizeof(RowGemmArgs):
            raise RuntimeError("libswinwnet_b200.so ABI mismatch (version or swn_rowgemm_args layout)")
        # the digest of the sources + flags is compiled into the library: a stale build must not run against newer
        # packing / Python code (the weight images and parameter structs are defined on both sides)
        from . import build as _build
        if os.path.isdir(_build.CSRC) and not VARIANT.startswith("x"):   # x*: scratch A/B builds (tools/), flags vary
            want, have = _build.digest(VARIANT), lib.swn_build_digest().decode()
            if want != have:
                raise RuntimeError(f"{LIB_PATH} is stale (built from other sources/flags: {have[:12]} != {want[:12]}); "
                                   "rebuild with `python __graft_entry__.py build`")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().swn_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
