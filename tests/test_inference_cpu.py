"""Point re-mapping and confusion-matrix helpers against the reference's per-scene numpy code."""
import numpy as np
import torch

from openpcseg_b200.segmentors import inference as I


def _reference_mapping(out, vox_b, inv, inv_b, num_points, softmax):
    """minkunet.py:441-452 restated: per scene boolean masks, then gather, then truncate."""
    res = []
    for idx in range(int(inv_b.max()) + 1):
        scene = out[vox_b == idx][inv[inv_b == idx]]
        scene = torch.from_numpy(scene).softmax(1).numpy() if softmax else scene.argmax(1)
        res.append(scene[: num_points[idx]])
    return res


def test_map_to_points_equals_per_scene_masks():
    rng = np.random.default_rng(0)
    n_vox, n_pts = [700, 1200, 50], [1500, 2600, 90]
    vox_b = np.concatenate([np.full(v, b) for b, v in enumerate(n_vox)])
    inv = np.concatenate([rng.integers(0, v, p) for v, p in zip(n_vox, n_pts)])
    inv_b = np.concatenate([np.full(p, b) for b, p in enumerate(n_pts)])
    out = rng.normal(size=(sum(n_vox), 20)).astype(np.float32)
    num_points = [1500, 2000, 90]
    for softmax in (False, True):
        ref = _reference_mapping(out, vox_b, inv, inv_b, num_points, softmax)
        got = I.map_to_points(torch.from_numpy(out), torch.from_numpy(vox_b), torch.from_numpy(inv),
                              torch.from_numpy(inv_b), num_points, "softmax" if softmax else "argmax")
        assert len(got) == 3
        for g, r in zip(got, ref):
            if softmax:
                np.testing.assert_allclose(g.numpy(), r, rtol=1e-6, atol=1e-7)
            else:
                np.testing.assert_array_equal(g.numpy(), r)
    # shuffled point order (not straight from sparse_collate) still groups by scan
    perm = rng.permutation(len(inv))
    got = I.map_to_points(torch.from_numpy(out), torch.from_numpy(vox_b), torch.from_numpy(inv[perm]),
                          torch.from_numpy(inv_b[perm]), None, "logits")
    for b in range(3):
        np.testing.assert_array_equal(got[b].numpy(), out[vox_b == b][inv[perm][inv_b[perm] == b]])


def test_confusion_matrix_and_iou():
    rng = np.random.default_rng(1)
    pred, label = rng.integers(0, 20, 5000), rng.integers(0, 20, 5000)
    unique_label = np.arange(19)                                     # classes 1..19 evaluated
    n = int(unique_label.max()) + 2
    keep = (label >= 0) & (label < n)
    ref = np.bincount(n * label[keep] + pred[keep], minlength=n * n)[: n * n].reshape(n, n)
    ref = ref[unique_label + 1, :][:, unique_label + 1]
    got = I.fast_hist_crop(torch.from_numpy(pred), torch.from_numpy(label), torch.from_numpy(unique_label))
    np.testing.assert_array_equal(got.numpy(), ref)
    iou_ref = np.diag(ref) / (ref.sum(1) + ref.sum(0) - np.diag(ref) + 1e-9)
    np.testing.assert_allclose(I.per_class_iu(got).numpy(), iou_ref, rtol=1e-12)


def test_tta_vote():
    torch.manual_seed(0)
    votes = [torch.randn(100, 20).softmax(1) for _ in range(10)]
    assert torch.equal(I.tta_vote(votes), torch.stack(votes).sum(0).argmax(1))
