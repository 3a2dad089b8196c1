"""Import alias: ``import pcd_b200`` == the hyphenated package directory
``a-multimodal-diffusion-based-model-for-point-cloud-completion_b200`` (whose name is not a
valid Python identifier).  Sub-modules are aliased too, so ``from pcd_b200.ops import x``
resolves to the same module objects."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_NAME = "a-multimodal-diffusion-based-model-for-point-cloud-completion_b200"
_pkg = importlib.import_module(_NAME)
for _k, _v in list(sys.modules.items()):
    if _k.startswith(_NAME + "."):
        sys.modules[__name__ + _k[len(_NAME):]] = _v
sys.modules[__name__] = _pkg
