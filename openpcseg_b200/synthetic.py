"""Deterministic synthetic SemanticKITTI-shaped scans.

HDL-64E-like: 64 elevation angles linspace(+2 deg, -24.8 deg) x 1875 azimuths = 120 000
rays; range = min(ground plane 1.73 m below the sensor, one wall per 10-degree sector at
U(8, 45) m, 80 m) + N(0, 0.02 m); intensity U(0, 1).  Pre-processing mirrors the voxel
dataset of the reference (pcseg/data/dataset/semantickitti/semantickitti_voxel.py:112-129):
``pc = round(xyz / voxel)``, shift to non-negative, keep one point per voxel.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .torchsparse.utils.quantize import sparse_quantize

__all__ = ["make_scan", "make_batch"]


def make_scan(seed: int = 0, voxel_size: float = 0.05, n_beams: int = 64, n_azimuth: int = 1875,
              num_class: int = 20) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))[:, None]
    azim = np.linspace(-np.pi, np.pi, n_azimuth, endpoint=False)[None, :]
    sector = ((azim + np.pi) / np.deg2rad(10.0)).astype(np.int64) % 36
    wall = rng.uniform(8.0, 45.0, size=36)[sector]                      # [1, A] horizontal distance
    with np.errstate(divide="ignore"):
        r_ground = np.where(np.sin(elev) < 0, 1.73 / -np.sin(elev), np.inf)
    r = np.minimum(np.minimum(r_ground, wall / np.cos(elev)), 80.0)
    r = r + rng.normal(0.0, 0.02, size=r.shape)
    xyz = np.stack([r * np.cos(elev) * np.cos(azim), r * np.cos(elev) * np.sin(azim),
                    r * np.sin(elev)], -1).reshape(-1, 3).astype(np.float32)
    intensity = rng.uniform(0.0, 1.0, size=(xyz.shape[0], 1)).astype(np.float32)
    pc = np.round(xyz / voxel_size).astype(np.int32)
    pc -= pc.min(0, keepdims=True)
    _, keep = sparse_quantize(pc, 1, return_index=True)
    feats = np.concatenate([xyz, intensity], 1)[keep]
    labels = rng.integers(0, num_class, size=keep.shape[0]).astype(np.int64)
    return {"coords": pc[keep].astype(np.int32), "feats": feats.astype(np.float32), "labels": labels,
            "n_raw_points": int(xyz.shape[0])}


def make_batch(seeds: List[int], **kw) -> Dict[str, np.ndarray]:
    """Collate scans: coords int32 [N, 4] = (x, y, z, batch), feats fp32 [N, 4], labels int64 [N]."""
    scans = [make_scan(s, **kw) for s in seeds]
    coords = np.concatenate([np.concatenate([s["coords"], np.full((len(s["coords"]), 1), b, np.int32)], 1)
                             for b, s in enumerate(scans)])
    return {"coords": coords, "feats": np.concatenate([s["feats"] for s in scans]),
            "labels": np.concatenate([s["labels"] for s in scans]),
            "n_scans": len(scans), "n_raw_points": sum(s["n_raw_points"] for s in scans)}
