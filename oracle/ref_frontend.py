"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's per-scan preprocessing.

Checks ``openpcseg_b200/frontend.py``.  Only tests may import this module.  Each
function follows the reference lines cited.  Pinning: the reference has no tests for its datasets, so
tests/golden/frontend.npz freezes outputs of the reference's OWN functions run in the build container
(sparse_quantize, cart2polar, voxelize_with_label, get_range_image; generator
tests/golden/make_golden_frontend.py) and tests/test_frontend_cpu.py checks these restatements and the product
against them.
"""
import numpy as np

from openpcseg_b200.torchsparse.utils.quantize import sparse_quantize   # host path = TS/utils/quantize.py:24-46


def voxel_scan_ref(point, point_label, voxel_size):
    """pcseg/data/dataset/semantickitti/semantickitti_voxel.py:112-141 (no augmentation, no drop)."""
    pc_ = np.round(point[:, :3] / voxel_size).astype(np.int32)
    pc_ -= pc_.min(0, keepdims=1)
    _, inds, inverse_map = sparse_quantize(pc_, return_index=True, return_inverse=True)
    return {"pc": pc_[inds], "feat": point[inds], "labels": point_label[inds], "pc_all": pc_,
            "inverse_map": inverse_map}


def voxelize_with_label_ref(point_coords, point_labels, num_classes):
    """semantickitti_cylinder.py:32-45: per-cell label histogram in a python loop, label 67 skipped."""
    voxel_coords, inds, inverse_map = sparse_quantize(point_coords, return_index=True, return_inverse=True)
    counter = np.zeros([voxel_coords.shape[0], num_classes])
    for n in range(len(inverse_map)):
        if point_labels[n] != 67:
            counter[inverse_map[n]][point_labels[n]] += 1
    return voxel_coords, np.argmax(counter, axis=1), inds, inverse_map


def cylinder_scan_ref(point, point_label, grid_size, min_bound, max_bound, num_classes):
    """semantickitti_cylinder.py:17-22 (cart2polar) and :137-171 (get_single_sample, eval branch)."""
    xyz = point[:, :3]
    rho = np.sqrt(xyz[:, 0] ** 2 + xyz[:, 1] ** 2)
    phi = np.arctan2(xyz[:, 1], xyz[:, 0])
    xyz_pol = np.stack((rho, phi, xyz[:, 2]), axis=1)
    xyz_pol[:, 1] = xyz_pol[:, 1] / np.pi * 180.
    max_bound, min_bound = np.asarray(max_bound, dtype=np.float64), np.asarray(min_bound, dtype=np.float64)
    intervals = (max_bound - min_bound) / (np.asarray(grid_size) - 1)
    point_coord = np.floor((np.clip(xyz_pol, min_bound, max_bound) - min_bound) / intervals).astype(np.int64)
    voxel_coord, voxel_label, inds, inverse_map = voxelize_with_label_ref(point_coord, point_label, num_classes)
    voxel_centers = (voxel_coord.astype(np.float32) + 0.5) * intervals + min_bound
    voxel_feature = np.concatenate([voxel_centers, xyz_pol[inds], point[inds][:, :2], point[inds][:, 3:]], axis=1)
    point_centers = (point_coord.astype(np.float32) + 0.5) * intervals + min_bound
    point_feature = np.concatenate([point_centers, xyz_pol, point[:, :2], point[:, 3:]], axis=1)
    return {"point_feature": point_feature.astype(np.float32), "point_coord": point_coord.astype(np.float32),
            "voxel_feature": voxel_feature.astype(np.float32), "voxel_coord": voxel_coord,
            "voxel_label": voxel_label, "inverse_map": inverse_map}


def range_projection_ref(points, yaw_offset, hw=(64, 2048)):
    """semantickitti_fusion.py:64-114 with the random cut passed in; the cv2 resize is the identity
    (INIT_HW == UP_HW)."""
    h, w = hw
    depth = np.linalg.norm(points[:, 0:3], 2, axis=1)
    yaw = np.arctan2(points[:, 1], -points[:, 0]) + yaw_offset
    yaw = yaw % (2 * np.pi) - np.pi
    proj_x = 0.5 * (yaw / np.pi + 1.0) * (w - 1)
    proj_x = np.round(proj_x).astype(np.int32)
    proj_y = np.round(points[:, 4]).astype(np.int32)
    proj_range = np.zeros((h, w))
    proj_range[proj_y, proj_x] = 1.0 / depth
    proj_refl = np.zeros((h, w))
    proj_refl[proj_y, proj_x] = points[:, 3]
    proj_xyz = np.zeros((h, w, 3))
    proj_xyz[proj_y, proj_x] = points[:, :3]
    px = 2.0 * (proj_x / (w - 1) - 0.5)
    py = 2.0 * (proj_y / (h - 1) - 0.5)
    image = np.concatenate([(25 * (proj_range - 0.4))[np.newaxis], (20 * (proj_refl - 0.5))[np.newaxis],
                            proj_xyz.transpose(2, 0, 1)]).astype(np.float32)
    return image, np.hstack([px.reshape(-1, 1), py.reshape(-1, 1)])
