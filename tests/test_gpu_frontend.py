"""Device data front-end against the reference-order host path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scan(n, seed, spread=20.0):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-spread, spread, size=(n, 3))
    pts[: n // 10] = pts[n // 2: n // 2 + n // 10]            # exact duplicates
    return pts


@pytest.mark.parametrize("n,voxel", [(1, 0.5), (257, 0.05), (20000, 0.2), (120000, 0.05), (50000, (0.1, 0.2, 0.4))])
def test_sparse_quantize_device_matches_host(n, voxel):
    from openpcseg_b200.torchsparse.utils.quantize import sparse_quantize

    pts = _scan(n, seed=n)
    vox, idx, inv = sparse_quantize(pts, voxel, return_index=True, return_inverse=True)
    d = torch.from_numpy(pts).cuda()
    gv, gi, gn = sparse_quantize(d, voxel, return_index=True, return_inverse=True)
    assert gv.dtype == torch.int32 and gv.is_cuda
    np.testing.assert_array_equal(gv.cpu().numpy(), vox)
    np.testing.assert_array_equal(gi.cpu().numpy(), idx)
    np.testing.assert_array_equal(gn.cpu().numpy(), inv)
    only = sparse_quantize(d, voxel)
    np.testing.assert_array_equal(only.cpu().numpy(), vox)


def test_sparse_quantize_device_empty_and_collate():
    from openpcseg_b200.torchsparse import SparseTensor
    from openpcseg_b200.torchsparse.utils.collate import sparse_collate
    from openpcseg_b200.torchsparse.utils.quantize import sparse_quantize

    e = sparse_quantize(torch.empty(0, 3, device="cuda"), 0.1, return_index=True)
    assert e[0].shape == (0, 3) and e[1].numel() == 0
    scans = []
    for b in range(3):
        d = torch.from_numpy(_scan(1000 + b, seed=b)).cuda()
        v, i = sparse_quantize(d, 0.1, return_index=True)
        scans.append(SparseTensor(d[i].float(), v))
    batch = sparse_collate(scans)
    assert batch.coords.is_cuda and batch.coords.shape[1] == 4
    assert batch.coords[:, 3].unique().tolist() == [0, 1, 2]
    assert batch.feats.shape[0] == sum(s.coords.shape[0] for s in scans)
