"""nn.utils: fapply and get_kernel_offsets (reference: TS/nn/utils/apply.py:10-16,
TS/nn/utils/kernel.py:11-32)."""
from functools import lru_cache
from typing import Callable

import numpy as np
import torch

from ..tensor import SparseTensor
from ..utils import make_ntuple

__all__ = ["fapply", "get_kernel_offsets"]


def fapply(input: SparseTensor, fn: Callable[..., torch.Tensor], *args, **kwargs) -> SparseTensor:
    return input._like(fn(input.feats, *args, **kwargs))


@lru_cache(maxsize=256)
def _offsets_np(size, stride, dilation) -> np.ndarray:
    # per-axis taps: size 3 -> (-1, 0, 1), size 2 -> (0, 1), scaled by stride * dilation
    taps = [np.arange((-size[a]) // 2 + 1, size[a] // 2 + 1) * stride[a] * dilation[a]
            for a in range(3)]
    if int(np.prod(size)) % 2 == 1:          # odd volume: x fastest (MinkowskiEngine layout)
        zz, yy, xx = np.meshgrid(taps[2], taps[1], taps[0], indexing="ij")
    else:                                    # even volume: z fastest
        xx, yy, zz = np.meshgrid(taps[0], taps[1], taps[2], indexing="ij")
    return np.stack([xx.ravel(), yy.ravel(), zz.ravel()], 1).astype(np.int32)


def get_kernel_offsets(size, stride=1, dilation=1, device="cpu") -> torch.Tensor:
    """int32 [K, 3] offsets; the row order defines which weight slice W[k] an offset uses."""
    size, stride, dilation = (make_ntuple(size, ndim=3), make_ntuple(stride, ndim=3),
                              make_ntuple(dilation, ndim=3))
    return torch.from_numpy(_offsets_np(size, stride, dilation).copy()).to(device)
