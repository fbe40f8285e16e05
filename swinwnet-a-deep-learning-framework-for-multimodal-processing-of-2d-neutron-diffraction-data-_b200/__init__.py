"""SwinWNet forward hot path on NVIDIA B200 (sm_100a): drop-in ``SwinWNet`` / ``SwinUNet`` / ``SwinUNetSR``
modules and the ``SwinWNetInference`` pipeline, lowered onto hand-written CUDA kernels behind a C ABI
(``include/swinwnet_b200.h`` -> ``libswinwnet_b200.so``).  Import name: ``swinwnet_b200`` (see the
repo-root ``swinwnet_b200.py`` loader; the directory name itself is not a valid Python identifier)."""
from .model import SwinWNet, SwinUNet, SwinUNetSR  # noqa: F401
from .pipeline import SwinWNetInference  # noqa: F401
from . import ops, packing, _lib, checkpoint, physics, dist, train, autograd, torch_ref  # noqa: F401

__all__ = ["SwinWNet", "SwinUNet", "SwinUNetSR", "SwinWNetInference", "ops", "packing", "checkpoint", "physics"]
