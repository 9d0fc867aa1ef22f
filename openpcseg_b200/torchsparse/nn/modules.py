"""spnn modules (reference: TS/nn/modules/conv.py:15-72, norm.py:10-41, activation.py:8-19).

Parameter layout contract kept for checkpoint compatibility: ``Conv3d.kernel`` is
``[K, C_in, C_out]`` (``[C_in, C_out]`` when K == 1), optional ``bias [C_out]``.
"""
import math

import numpy as np
import torch
from torch import nn

from ..tensor import SparseTensor
from ..utils import make_ntuple
from . import functional as F
from .utils import fapply

__all__ = ["Conv3d", "BatchNorm", "GroupNorm", "ReLU", "LeakyReLU"]


class Conv3d(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size=3, stride=1, dilation: int = 1,
                 bias: bool = False, transposed: bool = False) -> None:
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = make_ntuple(kernel_size, ndim=3)
        self.stride = make_ntuple(stride, ndim=3)
        self.dilation = dilation
        self.transposed = transposed
        self.kernel_volume = int(np.prod(self.kernel_size))
        shape = ((self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1
                 else (in_channels, out_channels))
        self.kernel = nn.Parameter(torch.zeros(*shape))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self) -> None:
        fan = (self.out_channels if self.transposed else self.in_channels) * self.kernel_volume
        bound = 1.0 / math.sqrt(fan)
        with torch.no_grad():
            self.kernel.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def extra_repr(self) -> str:
        parts = [f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}"]
        if any(s != 1 for s in self.stride):
            parts.append(f"stride={self.stride}")
        if self.dilation != 1:
            parts.append(f"dilation={self.dilation}")
        if self.bias is None:
            parts.append("bias=False")
        if self.transposed:
            parts.append("transposed=True")
        return ", ".join(parts)

    def forward(self, input: SparseTensor, bn_sums=None) -> SparseTensor:
        return F.conv3d(input, self.kernel, kernel_size=self.kernel_size, bias=self.bias,
                        stride=self.stride, dilation=self.dilation, transposed=self.transposed,
                        bn_sums=bn_sums)


class BatchNorm(nn.BatchNorm1d):
    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)


class GroupNorm(nn.GroupNorm):
    """Per-scan group norm over the rows of each batch index."""

    def forward(self, input: SparseTensor) -> SparseTensor:
        feats, batch = input.feats, input.coords[:, -1]
        out = torch.zeros_like(feats)
        for b in range(int(batch.max().item()) + 1):
            rows = batch == b
            chunk = feats[rows].t().unsqueeze(0)                     # [1, C, n_b]
            out[rows] = super().forward(chunk).squeeze(0).t()
        return input._like(out)


class ReLU(nn.ReLU):
    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)


class LeakyReLU(nn.LeakyReLU):
    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)
