"""Minimal ``torch_scatter`` stand-in for the one function Cylinder3D uses
(`torch_scatter.scatter_max`, tools/utils/common/seg_utils.py:178-180,
pcseg/model/segmentor/voxel/cylinder3d/cylinder_ts.py:35), on the libb2s kernel.
``openpcseg_b200.install_as_torchsparse()`` registers it as ``torch_scatter`` when the real
package is absent."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import backend as B

__all__ = ["scatter_max"]


class _ScatterMax(Function):
    @staticmethod
    def forward(ctx, src, index, dim_size):
        out, arg = B.scatter_max(src, index, dim_size)
        ctx.save_for_backward(arg)
        ctx.n = src.shape[0]
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        n, c = ctx.n, arg.shape[1]
        grad = grad_out.new_zeros((n + 1, c))                    # row n collects the empty voxels
        grad.scatter_(0, arg, grad_out.contiguous())
        return grad[:n], None, None


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim: int = 0, out=None,
                dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    assert dim == 0 and out is None and src.ndim == 2, "only the [N, C] / dim=0 form Cylinder3D uses"
    if index.ndim == 2:                                          # broadcast form index[:, None].expand
        index = index[:, 0]
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    return _ScatterMax.apply(src, index, int(dim_size))
