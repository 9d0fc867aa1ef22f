"""torchsparse-compatible operator surface on top of the B200-native backend.

Same names and semantics as the torchsparse 1.4.0 python package the reference
bundles (TS/__init__.py:1-3), so ``pcseg/model/**`` can run unmodified after
``openpcseg_b200.install_as_torchsparse()`` aliases this package as
``torchsparse``.  All compute goes through libb2s (CUDA, sm_100a).
"""
from .operators import cat
from .tensor import PointTensor, SparseTensor

__version__ = "1.4.0+b200"
__all__ = ["SparseTensor", "PointTensor", "cat", "__version__"]

from . import nn, utils  # noqa: E402,F401  (import order: tensor first, like the reference)
