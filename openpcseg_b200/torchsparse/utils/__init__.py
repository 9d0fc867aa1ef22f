"""Host-side helpers (reference: TS/utils/utils.py:9-20)."""
from typing import Tuple

import torch

__all__ = ["make_ntuple"]


def make_ntuple(x, ndim: int) -> Tuple[int, ...]:
    if isinstance(x, int):
        return (x,) * ndim
    if isinstance(x, torch.Tensor):
        x = x.view(-1).cpu().tolist()
    if isinstance(x, list):
        x = tuple(x)
    assert isinstance(x, tuple) and len(x) == ndim, x
    return x
