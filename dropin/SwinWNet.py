"""Drop-in for the reference's ``SwinWNet`` module: put this directory on sys.path instead of the
reference checkout and ``from SwinWNet import SwinWNet, SwinUNet, SwinUNetSR`` resolves to the B200 path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swinwnet_b200  # noqa: E402
from swinwnet_b200.model import *  # noqa: E402,F401,F403
from swinwnet_b200.model import SwinWNet, SwinUNet, SwinUNetSR  # noqa: E402,F401
