"""Import shim: exposes the package directory
``swinwnet-a-deep-learning-framework-for-multimodal-processing-of-2d-neutron-diffraction-data-_b200/``
(not a valid identifier) as the module ``swinwnet_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "swinwnet-a-deep-learning-framework-for-multimodal-processing-of-2d-neutron-diffraction-data-_b200")
_spec = importlib.util.spec_from_file_location("swinwnet_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["swinwnet_b200"] = _mod
_spec.loader.exec_module(_mod)
