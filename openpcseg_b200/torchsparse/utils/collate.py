"""Batch collation (reference: TS/utils/collate.py:11-59): append the batch index as the
4th coordinate column and concatenate scans."""
from typing import Any, List

import numpy as np
import torch

from ..tensor import SparseTensor

__all__ = ["sparse_collate", "sparse_collate_fn"]


def _as_tensor(x):
    return torch.as_tensor(x) if isinstance(x, np.ndarray) else x


def sparse_collate(inputs: List[SparseTensor]) -> SparseTensor:
    stride = inputs[0].stride
    coords, feats = [], []
    for b, x in enumerate(inputs):
        x.coords, x.feats = _as_tensor(x.coords), _as_tensor(x.feats)
        assert isinstance(x.coords, torch.Tensor) and isinstance(x.feats, torch.Tensor)
        assert x.stride == stride, (x.stride, stride)
        bcol = x.coords.new_full((x.coords.shape[0], 1), b, dtype=torch.int)
        coords.append(torch.cat((x.coords, bcol), dim=1))
        feats.append(x.feats)
    return SparseTensor(torch.cat(feats, 0), torch.cat(coords, 0), stride)


def sparse_collate_fn(inputs: List[Any]) -> Any:
    if not isinstance(inputs[0], dict):
        return inputs
    out = {}
    for name, first in inputs[0].items():
        column = [item[name] for item in inputs]
        if isinstance(first, dict):
            out[name] = sparse_collate_fn(column)
        elif isinstance(first, np.ndarray):
            out[name] = torch.stack([torch.as_tensor(v) for v in column], 0)
        elif isinstance(first, torch.Tensor):
            out[name] = torch.stack(column, 0)
        elif isinstance(first, SparseTensor):
            out[name] = sparse_collate(column)
        else:
            out[name] = column
    return out
