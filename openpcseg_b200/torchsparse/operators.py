"""Tensor-level operators (reference: TS/operators.py:10-17)."""
from typing import List

import torch

from .tensor import SparseTensor

__all__ = ["cat"]


def cat(inputs: List[SparseTensor]) -> SparseTensor:
    """Channel-concatenate sparse tensors living on the same coordinates."""
    head = inputs[0]
    return head._like(torch.cat([t.feats for t in inputs], dim=1))
