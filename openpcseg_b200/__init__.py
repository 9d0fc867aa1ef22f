"""openpcseg_b200: B200-native (sm_100a) sparse-voxel convolution backend that drops in
under OpenPCSeg's voxel / fusion segmentors behind the torchsparse operator surface.

    import openpcseg_b200
    openpcseg_b200.install_as_torchsparse()   # `import torchsparse` now resolves here
"""
import sys

__all__ = ["install_as_torchsparse"]


def install_as_torchsparse() -> None:
    """Alias ``openpcseg_b200.torchsparse`` (and submodules) as ``torchsparse`` so that
    reference model code (`import torchsparse.nn as spnn`, ...) runs unmodified."""
    import importlib

    pkg = importlib.import_module("openpcseg_b200.torchsparse")
    prefix = "openpcseg_b200.torchsparse"
    for name in ["", ".nn", ".nn.functional", ".nn.utils", ".nn.modules", ".utils", ".utils.collate",
                 ".utils.quantize", ".tensor", ".operators"]:
        sys.modules["torchsparse" + name] = importlib.import_module(prefix + name)
    from . import backend
    sys.modules["torchsparse.backend"] = backend
    pkg.backend = backend
    # RPVNet's range-image ops (`import range_utils.nn.functional as rnf`, rpvnet.py:26)
    for name in ["", ".nn", ".nn.functional"]:
        sys.modules["range_utils" + name] = importlib.import_module("openpcseg_b200.range_utils" + name)
    try:                                  # Cylinder3D's scatter_max when torch_scatter is not installed
        importlib.import_module("torch_scatter")
    except ImportError:
        sys.modules["torch_scatter"] = importlib.import_module("openpcseg_b200.torch_scatter")
