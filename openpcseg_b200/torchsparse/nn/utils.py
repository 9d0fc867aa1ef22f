"""nn.utils: fapply and get_kernel_offsets (reference: TS/nn/utils/apply.py:10-16,
TS/nn/utils/kernel.py:11-32)."""
from functools import lru_cache
from typing import Callable

import numpy as np
import torch

from ..tensor import SparseTensor
from ..utils import make_ntuple

__all__ = ["fapply", "get_kernel_offsets"]


_BN_FORWARDS = (torch.nn.modules.batchnorm._BatchNorm.forward, torch.nn.SyncBatchNorm.forward)
_FUSE_FAPPLY_BN = __import__("os").environ.get("B2S_FAPPLY_BN", "1") != "0"


def fapply(input: SparseTensor, fn: Callable[..., torch.Tensor], *args, **kwargs) -> SparseTensor:
    """TS/nn/utils/apply.py:10-16: ``fn`` on the feature rows, maps re-attached.

    The reference's segmentors wrap their norms as ``fapply(input, super().forward)`` on a BatchNorm1d /
    SyncBatchNorm subclass (minkunet.py:23-29, spvcnn.py, rpvnet.py:256-263, cylinder_ts.py).  When ``fn`` is
    exactly that bound method and the rows live on the GPU, the same training-mode batch norm is computed by
    this backend's fused kernels (statistics in fp64, one apply pass; synchronised variant = one fp64 all-reduce
    per direction) instead of torch's generic channels-last kernels, which cost 38 ms of a 108 ms MinkUNet-34 step
    (profiles/r2_kernels_minkunet34_ref.txt).  Eval mode, unsupported widths and every other ``fn`` take the
    plain path.  B2S_FAPPLY_BN=0 switches the interception off."""
    bn = getattr(fn, "__self__", None)
    if (_FUSE_FAPPLY_BN and not args and not kwargs and isinstance(bn, torch.nn.modules.batchnorm._BatchNorm)
            and getattr(fn, "__func__", None) in _BN_FORWARDS and input.feats.is_cuda and input.feats.ndim == 2):
        from .functional import batch_norm_act
        return input._like(batch_norm_act(input.feats, bn, relu=False))
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


_OFFSETS_DEV = {}


def kernel_offsets_cached(size, stride=1, dilation=1, device="cpu") -> torch.Tensor:
    """``get_kernel_offsets`` for this package's own READ-ONLY use: one device copy per (size, stride, dilation,
    device) instead of a pageable host-to-device copy (a blocking call) per kernel map and per trilinear map."""
    key = (make_ntuple(size, ndim=3), make_ntuple(stride, ndim=3), make_ntuple(dilation, ndim=3), torch.device(device))
    hit = _OFFSETS_DEV.get(key)
    if hit is None:
        hit = _OFFSETS_DEV[key] = get_kernel_offsets(size, stride, dilation, device)
    return hit
