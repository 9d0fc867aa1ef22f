"""Evaluation-side glue of the voxel segmentors as tensor programs.

The reference maps voxel predictions back to the raw points of every scan through per-scene boolean
masks built on the host (``.cpu().numpy()`` three times per scene,
pcseg/model/segmentor/voxel/minkunet/minkunet.py:435-455) and accumulates the confusion matrix with
numpy (infer.py:35-52).  Scans are concatenated in batch order by ``sparse_collate``, so the voxel row of
a raw point is ``first_voxel_row_of_its_scan + inverse_map`` - one gather for the whole batch, no masks,
no host round trip until the caller asks for the result.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

__all__ = ["map_to_points", "tta_vote", "fast_hist", "fast_hist_crop", "per_class_iu"]


def map_to_points(out: torch.Tensor, voxel_batch: torch.Tensor, inverse: torch.Tensor,
                  point_batch: torch.Tensor, num_points: Optional[Sequence[int]] = None,
                  mode: str = "argmax") -> List[torch.Tensor]:
    """Per-scan point predictions from per-voxel logits.

    out ``[V, C]`` logits of the collated batch; voxel_batch ``[V]`` = ``x.C[:, -1]``; inverse ``[P]`` =
    ``inverse_map.F`` (row inside the scan); point_batch ``[P]`` = ``inverse_map.C[:, -1]``.
    mode: "argmax" (labels), "softmax" (probabilities, the TTA / return_logit branch) or "logits".
    ``num_points[b]`` truncates scan b like the reference (multi-frame inputs)."""
    n_scans = int(point_batch.max().item()) + 1 if point_batch.numel() else 0
    per_scan = torch.bincount(voxel_batch.long(), minlength=n_scans)
    first_row = torch.cumsum(per_scan, 0) - per_scan
    rows = first_row[point_batch.long()] + inverse.long()
    mapped = out.index_select(0, rows)
    if mode == "argmax":
        mapped = mapped.argmax(1)
    elif mode == "softmax":
        mapped = mapped.softmax(1)
    elif mode != "logits":
        raise ValueError(f"mode {mode!r}")
    pts_per_scan = torch.bincount(point_batch.long(), minlength=n_scans).tolist()
    # points of one scan are contiguous when the batch came from sparse_collate; otherwise select
    ordered = bool((point_batch[1:] >= point_batch[:-1]).all()) if point_batch.numel() > 1 else True
    res, start = [], 0
    for b, cnt in enumerate(pts_per_scan):
        piece = mapped[start:start + cnt] if ordered else mapped[point_batch == b]
        start += cnt
        if num_points is not None:
            piece = piece[: int(num_points[b])]
        res.append(piece)
    return res


def tta_vote(probabilities: Sequence[torch.Tensor]) -> torch.Tensor:
    """Labels from test-time-augmentation votes: arg max of the summed per-vote softmax outputs."""
    total = probabilities[0].clone()
    for p in probabilities[1:]:
        total += p
    return total.argmax(1)


def fast_hist(pred: torch.Tensor, label: torch.Tensor, n: int) -> torch.Tensor:
    """Confusion matrix ``[n, n]`` (rows = label) over points with 0 <= label < n (infer.py:35-40)."""
    keep = (label >= 0) & (label < n)
    flat = n * label[keep].long() + pred[keep].long()
    return torch.bincount(flat, minlength=n * n)[: n * n].view(n, n)


def fast_hist_crop(output: torch.Tensor, target: torch.Tensor, unique_label: torch.Tensor) -> torch.Tensor:
    """Confusion matrix restricted to the evaluated classes (infer.py:47-52): labels are shifted by one
    (0 = ignored) and ``unique_label`` holds the zero-based ids of the evaluated classes."""
    hist = fast_hist(output.flatten(), target.flatten(), int(unique_label.max().item()) + 2)
    sel = unique_label.long() + 1
    return hist.index_select(0, sel).index_select(1, sel)


def per_class_iu(hist: torch.Tensor) -> torch.Tensor:
    hist = hist.double()
    diag = torch.diagonal(hist)
    return diag / (hist.sum(1) + hist.sum(0) - diag + 1e-9)
