"""Data front-end of the voxel / cylinder / fusion segmentors as tensor programs.

The reference prepares every scan on CPU workers with numpy (``np.round``, ``sparse_quantize``, an O(N)
python loop for the cylinder majority labels, fancy-index scatter for the range image).  The same
results are produced here from torch ops that run wherever the input tensor lives; on a CUDA tensor the
unique/first-index/inverse step is the backend's device sort + hash table
(``torchsparse.utils.quantize.sparse_quantize_device``), so a raw scan can go H2D once and never come
back.  Augmentation (``aug_points``) is the caller's business - it is random and not part of the path.

reference:
  voxel     pcseg/data/dataset/semantickitti/semantickitti_voxel.py:112-141
  cylinder  pcseg/data/dataset/semantickitti/semantickitti_cylinder.py:17-45, 137-171
  range     pcseg/data/dataset/semantickitti/semantickitti_fusion.py:64-114 (projection only)
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .torchsparse import SparseTensor
from .torchsparse.utils.quantize import sparse_quantize

__all__ = ["quantize_rows", "voxel_scan", "voxelize_with_label", "cylinder_scan", "range_projection"]


def quantize_rows(coords: torch.Tensor):
    """(unique rows in ravel order, first index, inverse) of an integer ``[N, 3]`` tensor."""
    if coords.is_cuda:
        return sparse_quantize(coords, 1, return_index=True, return_inverse=True)
    vox, first, inverse = sparse_quantize(coords.numpy(), 1, return_index=True, return_inverse=True)
    return torch.from_numpy(vox), torch.from_numpy(first), torch.from_numpy(inverse)


def voxel_scan(points: torch.Tensor, labels: Optional[torch.Tensor], voxel_size: float) -> Dict[str, object]:
    """One scan of the voxel datasets: ``points [N, >=4] = (x, y, z, intensity, ...)`` float.

    Returns the reference's sample dict entries: ``lidar`` (one point per voxel, feats = the raw point
    row, coords = round(xyz / voxel) min-shifted), ``targets``, ``targets_mapped``, ``inverse_map``."""
    pc = torch.round(points[:, :3] / voxel_size).to(torch.int32)
    pc = pc - pc.min(dim=0, keepdim=True).values
    _, first, inverse = quantize_rows(pc)
    out = {"lidar": SparseTensor(points.index_select(0, first), pc.index_select(0, first)),
           "inverse_map": SparseTensor(inverse, pc),
           "num_points": int(points.shape[0])}
    if labels is not None:
        labels = labels.reshape(-1)
        out["targets"] = SparseTensor(labels.index_select(0, first), pc.index_select(0, first))
        out["targets_mapped"] = SparseTensor(labels, pc)
    return out


def voxelize_with_label(point_coords: torch.Tensor, point_labels: torch.Tensor, num_classes: int,
                        skip_label: int = 67):
    """Unique cells of integer ``point_coords`` + the majority label of every cell.

    The reference counts labels per cell in a python loop (label 67 is not counted) and takes
    ``np.argmax`` (first maximum); here the histogram is one scatter-add."""
    voxel_coords, first, inverse = quantize_rows(point_coords.to(torch.int32))
    labels = point_labels.reshape(-1).long()
    counted = labels != skip_label
    flat = inverse[counted] * num_classes + labels[counted]
    hist = torch.zeros(voxel_coords.shape[0] * num_classes, dtype=torch.int32, device=point_coords.device)
    hist.scatter_add_(0, flat, torch.ones_like(flat, dtype=torch.int32))
    hist = hist.view(-1, num_classes)
    # first maximum, like np.argmax: maximise (count, -class)
    score = hist.long() * num_classes + (num_classes - 1 - torch.arange(num_classes, device=hist.device))
    voxel_labels = score.argmax(dim=1)
    return voxel_coords, voxel_labels, first, inverse


def cylinder_scan(points: torch.Tensor, labels: torch.Tensor, grid_size: Sequence[int],
                  min_bound: Sequence[float], max_bound: Sequence[float], num_classes: int) -> Dict[str, torch.Tensor]:
    """Cylinder3D sample: polar cells (rho, phi in degrees, z), 9-d point / voxel features, labels."""
    dev, f64 = points.device, torch.float64
    # polar coordinates in the dtype of the scan (numpy keeps float32 here), everything after in float64
    xyz = points[:, :3]
    rho = torch.sqrt(xyz[:, 0] ** 2 + xyz[:, 1] ** 2)
    phi = torch.atan2(xyz[:, 1], xyz[:, 0]) / math.pi * 180.0
    pol = torch.stack((rho, phi, xyz[:, 2]), dim=1).to(f64)
    lo = torch.as_tensor(min_bound, dtype=f64, device=dev)
    hi = torch.as_tensor(max_bound, dtype=f64, device=dev)
    grid = torch.as_tensor(grid_size, dtype=f64, device=dev)
    intervals = (hi - lo) / (grid - 1)
    coord = torch.floor((torch.minimum(torch.maximum(pol, lo), hi) - lo) / intervals).long()
    v_coord, v_label, first, inverse = voxelize_with_label(coord, labels, num_classes)
    # the reference forms the centres in float32 ((coord.astype(float32) + 0.5) * intervals + min_bound
    # promotes to float64 because intervals is float64) and casts the features to float32 at the end
    v_center = (v_coord.to(torch.float32).to(f64) + 0.5) * intervals + lo
    p_center = (coord.to(torch.float32).to(f64) + 0.5) * intervals + lo
    pts = points.to(f64)
    v_feat = torch.cat([v_center, pol.index_select(0, first), pts.index_select(0, first)[:, :2],
                        pts.index_select(0, first)[:, 3:]], dim=1)
    p_feat = torch.cat([p_center, pol, pts[:, :2], pts[:, 3:]], dim=1)
    return {"point_feature": p_feat.float(), "point_coord": coord.float(), "point_label": labels.reshape(-1).long(),
            "voxel_feature": v_feat.float(), "voxel_coord": v_coord.long(), "voxel_label": v_label.long(),
            "inverse_map": inverse.long(), "num_points": int(points.shape[0])}


def range_projection(points: torch.Tensor, yaw_offset: float = 0.0, hw: Sequence[int] = (64, 2048)):
    """Spherical projection of ``points [N, 5] = (x, y, z, reflectivity, ring)`` (RPVNet).

    Returns ``(image [5, H, W] float32, pxpy [N, 2])``: channels 25 * (1/depth - 0.4), 20 * (refl - 0.5),
    xyz.  ``yaw_offset`` is the reference's random cut ``(rand - 0.5) * 2 pi``.  Where several points fall
    into one pixel numpy's fancy assignment keeps the LAST one in input order; that is reproduced with a
    scatter-max of the point index.  (The reference then resizes with cv2 to the same size: identity.)"""
    h, w = int(hw[0]), int(hw[1])
    p = points                                       # angles / depth in the dtype of the scan, like numpy
    depth = torch.linalg.norm(p[:, :3], dim=1)
    yaw = torch.atan2(p[:, 1], -p[:, 0]) + yaw_offset
    yaw = torch.remainder(yaw, 2 * math.pi) - math.pi
    proj_x = torch.round(0.5 * (yaw / math.pi + 1.0) * (w - 1)).long()
    proj_y = torch.round(p[:, 4]).long()
    assert int(proj_y.max()) <= h - 1, "ring id exceeds the image height"
    pix = proj_y * w + proj_x
    last = torch.full((h * w,), -1, dtype=torch.long, device=points.device)
    last.scatter_reduce_(0, pix, torch.arange(p.shape[0], device=points.device), reduce="amax", include_self=True)
    hit = last >= 0
    src = last.clamp(min=0)
    f64 = torch.float64                              # the image planes are float64 in the reference
    zero = torch.zeros(h * w, dtype=f64, device=points.device)
    img = torch.empty(5, h * w, dtype=f64, device=points.device)
    img[0] = 25.0 * (torch.where(hit, (1.0 / depth[src]).to(f64), zero) - 0.4)
    img[1] = 20.0 * (torch.where(hit, p[src, 3].to(f64), zero) - 0.5)
    for c in range(3):
        img[2 + c] = torch.where(hit, p[src, c].to(f64), zero)
    px = 2.0 * (proj_x.to(f64) / (w - 1) - 0.5)
    py = 2.0 * (proj_y.to(f64) / (h - 1) - 0.5)
    return img.view(5, h, w).float(), torch.stack((px, py), dim=1)
