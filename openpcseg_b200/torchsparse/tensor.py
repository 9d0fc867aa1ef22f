"""SparseTensor / PointTensor containers (reference: TS/tensor.py:10-105).

``cmaps`` (stride -> coords) and ``kmaps`` ((stride, kernel, stride, dilation) ->
kernel map) are shared *by reference* between every tensor derived from one
input - that sharing is what amortises map construction over the ~12 convs of
a UNet level, so derived tensors must alias the dicts, never copy them.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch

from .utils import make_ntuple

__all__ = ["SparseTensor", "PointTensor"]


class SparseTensor:
    __slots__ = ("feats", "coords", "stride", "cmaps", "kmaps")

    def __init__(self, feats: torch.Tensor, coords: torch.Tensor, stride=1) -> None:
        self.feats = feats
        self.coords = coords
        self.stride = make_ntuple(stride, ndim=3)
        self.cmaps: Dict[Tuple[int, ...], torch.Tensor] = {}
        self.kmaps: Dict[Tuple[Any, ...], Any] = {}

    # short aliases used throughout pcseg (x.F / x.C / x.s)
    F = property(lambda self: self.feats, lambda self, v: setattr(self, "feats", v))
    C = property(lambda self: self.coords, lambda self, v: setattr(self, "coords", v))
    s = property(lambda self: self.stride,
                 lambda self, v: setattr(self, "stride", make_ntuple(v, ndim=3)))

    def _like(self, feats: torch.Tensor) -> "SparseTensor":
        out = SparseTensor(feats, self.coords, self.stride)
        out.cmaps, out.kmaps = self.cmaps, self.kmaps
        return out

    def _move(self, fn) -> "SparseTensor":
        self.coords, self.feats = fn(self.coords), fn(self.feats)
        return self

    def cpu(self):
        return self._move(lambda t: t.cpu())

    def cuda(self):
        return self._move(lambda t: t.cuda())

    def detach(self):
        return self._move(lambda t: t.detach())

    def to(self, device, non_blocking: bool = True):
        return self._move(lambda t: t.to(device, non_blocking=non_blocking))

    def __add__(self, other: "SparseTensor") -> "SparseTensor":
        return self._like(self.feats + other.feats)


class PointTensor:
    def __init__(self, feats, coords, idx_query: Optional[dict] = None,
                 weights: Optional[dict] = None) -> None:
        self.F = feats
        self.C = coords
        self.idx_query = {} if idx_query is None else idx_query
        self.weights = {} if weights is None else weights
        self.additional_features = {"idx_query": {}, "counts": {}}

    def _move(self, fn) -> "PointTensor":
        self.F, self.C = fn(self.F), fn(self.C)
        return self

    def cuda(self):
        return self._move(lambda t: t.cuda())

    def detach(self):
        return self._move(lambda t: t.detach())

    def to(self, device, non_blocking: bool = True):
        return self._move(lambda t: t.to(device, non_blocking=non_blocking))

    def __add__(self, other: "PointTensor") -> "PointTensor":
        out = PointTensor(self.F + other.F, self.C, self.idx_query, self.weights)
        out.additional_features = self.additional_features
        return out
