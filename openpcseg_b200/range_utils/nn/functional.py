"""range_utils.nn.functional: ``map_count`` and ``denselize`` with the reference's signatures,
shapes and autograd contract (range_lib/range_utils/nn/functional/map_count.py:9-29,
denselize.py:7-37; bindings range_utils/src/rangelib_bindings_gpu.cpp:7-12).

  map_count(pxpy int32 [N,3]=(batch,px,py), max_bs, h, w) -> int32 [max_bs, h, w]
  denselize(feat [N,C], count_map int32 [B,H,W], pxpy int32 [N,3]) -> fp32 [B,C,H,W]
      out[b,:,py,px] = mean of the rows that fall into the pixel; backward = gather / count
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import backend as B

__all__ = ["map_count", "denselize"]


def map_count(pxpy: torch.Tensor, max_bs: int, h: int, w: int) -> torch.Tensor:
    """Points per range-image pixel (map_count.cpp:5-17: zeros [max_bs, h, w] int32 + atomic count)."""
    assert pxpy.dtype == torch.int32 and pxpy.ndim == 2 and pxpy.shape[1] == 3, (pxpy.dtype, pxpy.shape)
    return B.map_count(pxpy, int(max_bs), int(h), int(w)).view(int(max_bs), int(h), int(w))


class _Denselize(Function):
    @staticmethod
    def forward(ctx, feat, count_map, pxpy):
        count_map, pxpy = count_map.int().contiguous(), pxpy.int().contiguous()
        b, (h, w) = count_map.shape[0], count_map.shape[-2:]
        ctx.aux = (count_map, pxpy)
        ctx.in_dtype = feat.dtype
        return B.denselize_forward(feat, count_map.view(b, 1, h, w), pxpy)

    @staticmethod
    @once_differentiable
    def backward(ctx, top_grad):
        count_map, pxpy = ctx.aux
        b, (h, w) = count_map.shape[0], count_map.shape[-2:]
        g = B.denselize_backward(top_grad, count_map.view(b, 1, h, w), pxpy)
        return g.to(ctx.in_dtype), None, None


def denselize(feat: torch.Tensor, count_map: torch.Tensor, pxpy: torch.Tensor) -> torch.Tensor:
    """Scatter-mean of point rows into a [B, C, H, W] image (denselize.cpp:5-20, denselize_gpu.cu:5-19)."""
    return _Denselize.apply(feat, count_map, pxpy)
