"""Tensor-level operators (reference: TS/operators.py:10-17)."""
from typing import List

import torch

from .tensor import SparseTensor

__all__ = ["cat"]


def _slice_cols(grad: torch.Tensor, c0: int, c1: int) -> torch.Tensor:
    """grad[:, c0:c1] as a contiguous tensor.

    The stock narrow().contiguous() of a 2-byte dtype is a scalar strided copy (0.7 TB/s measured on
    the [1.5 M, 128] fp16 skip connections, profiles/r1_launches_final.txt); when the column window
    is 16-byte aligned the same copy is done on a 16-byte-element view, i.e. vectorised."""
    es = grad.element_size()
    if (grad.is_contiguous() and grad.dim() == 2 and (c0 * es) % 16 == 0 and (c1 * es) % 16 == 0
            and (grad.shape[1] * es) % 16 == 0 and grad.data_ptr() % 16 == 0 and grad.shape[0] > 0):
        wide = grad.view(torch.complex128)
        return wide[:, c0 * es // 16:c1 * es // 16].contiguous().view(grad.dtype)
    return grad[:, c0:c1].contiguous()


class _CatFeats(torch.autograd.Function):
    """torch.cat(dim=1) whose backward hands out contiguous, vector-copied column windows."""

    @staticmethod
    def forward(ctx, *feats):
        ctx.widths = [f.shape[1] for f in feats]
        return torch.cat(feats, dim=1)

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous()
        outs, c0 = [], 0
        for i, w in enumerate(ctx.widths):
            outs.append(_slice_cols(grad, c0, c0 + w) if ctx.needs_input_grad[i] else None)
            c0 += w
        return tuple(outs)


def cat(inputs: List[SparseTensor]) -> SparseTensor:
    """Channel-concatenate sparse tensors living on the same coordinates."""
    head = inputs[0]
    return head._like(_CatFeats.apply(*[t.feats for t in inputs]))
