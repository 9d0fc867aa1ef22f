"""Callers of the hot path needed to measure it end to end (MinkUNet + point<->voxel glue).
The model graph stays PyTorch; every sparse op goes through libb2s."""
from .minkunet import MinkUNet, minkunet34_config  # noqa: F401
from .point_voxel import initial_voxelize, point_to_voxel, voxel_to_point  # noqa: F401
