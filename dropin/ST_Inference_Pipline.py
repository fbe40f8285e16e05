"""Drop-in for the reference's ``ST_Inference_Pipline`` module (same class name and attributes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swinwnet_b200  # noqa: E402
from swinwnet_b200.pipeline import SwinWNetInference  # noqa: E402,F401
