"""CPU-side tests: the C-ABI library loads and exports every symbol include/b2s.h declares,
and the host logic of the torchsparse-compatible surface (no GPU compute calls)."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import ref_ops as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from openpcseg_b200 import _lib, build
    build.build()
    return _lib.lib()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b2s.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from openpcseg_b200 import _lib
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(built_lib, name), f"libb2s.so does not export {name}"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "ctypes table out of sync with include/b2s.h"
    assert built_lib.b2s_version() >= 100
    assert built_lib.b2s_last_error() == b"ok" or isinstance(built_lib.b2s_last_error(), bytes)


def test_size_queries_run_without_a_gpu(built_lib):
    assert built_lib.b2s_table_slots(1000) == 2048
    assert built_lib.b2s_table_slots(0) == 1024
    assert built_lib.b2s_table_bytes(1000) == 2048 * 12


def test_missing_gpu_is_loud():
    import openpcseg_b200.torchsparse.nn.functional as F
    from openpcseg_b200._lib import B2SError
    with pytest.raises(B2SError):
        F.sphash(torch.zeros((4, 4), dtype=torch.int32))
    with pytest.raises(B2SError):
        F.spcount(torch.zeros(4, dtype=torch.int32), 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "openpcseg_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_kernel_offsets_and_ntuple():
    from openpcseg_b200.torchsparse.nn.utils import get_kernel_offsets
    from openpcseg_b200.torchsparse.utils import make_ntuple
    assert make_ntuple(2, 3) == (2, 2, 2) and make_ntuple([1, 2, 3], 3) == (1, 2, 3)
    assert make_ntuple(torch.tensor([4, 5, 6]), 3) == (4, 5, 6)
    with pytest.raises(AssertionError):
        make_ntuple((1, 2), 3)
    off = get_kernel_offsets(3)
    assert off.dtype == torch.int32 and off.shape == (27, 3)
    assert off[0].tolist() == [-1, -1, -1] and off[1].tolist() == [0, -1, -1] and off[13].tolist() == [0, 0, 0]
    assert get_kernel_offsets(2).tolist() == [[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0],
                                              [1, 0, 1], [1, 1, 0], [1, 1, 1]]
    for ks, st, dil in [((1, 3, 3), 1, 1), ((3, 1, 3), 2, 1), (2, 4, 1), (3, (2, 2, 1), 2), (5, 1, 1)]:
        assert np.array_equal(get_kernel_offsets(ks, st, dil).numpy(), R.get_kernel_offsets(ks, st, dil))


def test_tensor_containers_share_maps():
    import openpcseg_b200.torchsparse as ts
    x = ts.SparseTensor(torch.randn(5, 3), torch.zeros(5, 4, dtype=torch.int32), 2)
    assert x.s == (2, 2, 2) and x.F is x.feats and x.C is x.coords
    x.cmaps[(2, 2, 2)] = x.coords
    y = x + x
    z = ts.cat([x, y])
    w = ts.nn.utils.fapply(x, torch.relu)
    for t in (y, z, w):
        assert t.cmaps is x.cmaps and t.kmaps is x.kmaps and t.stride == x.stride
    assert z.F.shape == (5, 6) and torch.equal(y.F, 2 * x.F)
    x.F = x.F[:, :2]
    assert x.feats.shape == (5, 2)
    p = ts.PointTensor(torch.randn(4, 2), torch.rand(4, 4))
    q = p + p
    assert q.idx_query is p.idx_query and q.additional_features is p.additional_features
    assert set(p.additional_features) == {"idx_query", "counts"}


def test_quantize_and_collate():
    from openpcseg_b200.torchsparse import SparseTensor
    from openpcseg_b200.torchsparse.utils.collate import sparse_collate_fn
    from openpcseg_b200.torchsparse.utils.quantize import sparse_quantize
    rng = np.random.default_rng(0)
    pts = rng.uniform(-3, 3, size=(500, 3))
    vox, idx, inv = sparse_quantize(pts, 0.5, return_index=True, return_inverse=True)
    fl = np.floor(pts / 0.5).astype(np.int32)
    assert np.array_equal(vox, fl[idx]) and np.array_equal(vox[inv], fl)
    assert len(np.unique(vox, axis=0)) == len(vox)
    items = [{"lidar": SparseTensor(np.ones((3, 2), np.float32), np.zeros((3, 3), np.int32)),
              "name": "a", "n": np.array([1])},
             {"lidar": SparseTensor(np.ones((2, 2), np.float32), np.ones((2, 3), np.int32)),
              "name": "b", "n": np.array([2])}]
    out = sparse_collate_fn(items)
    assert out["lidar"].coords.shape == (5, 4) and out["lidar"].coords[:, 3].tolist() == [0, 0, 0, 1, 1]
    assert out["name"] == ["a", "b"] and out["n"].shape == (2, 1)


def test_conv3d_module_parameter_layout():
    import openpcseg_b200.torchsparse.nn as spnn
    m = spnn.Conv3d(8, 16, 3)
    assert m.kernel.shape == (27, 8, 16) and m.bias is None
    assert spnn.Conv3d(8, 16, 1).kernel.shape == (8, 16)
    assert spnn.Conv3d(8, 16, (1, 3, 3), bias=True).kernel.shape == (9, 8, 16)
    bound = 1 / np.sqrt(8 * 27)
    assert float(m.kernel.abs().max()) <= bound + 1e-7
    t = spnn.Conv3d(8, 16, 2, stride=2, transposed=True)
    assert float(t.kernel.abs().max()) <= 1 / np.sqrt(16 * 8) + 1e-7
    assert "transposed=True" in repr(t)


def test_install_as_torchsparse():
    import sys
    import openpcseg_b200
    openpcseg_b200.install_as_torchsparse()
    import torchsparse
    import torchsparse.nn as spnn
    import torchsparse.nn.functional as F
    from torchsparse.nn.utils import fapply, get_kernel_offsets  # noqa: F401
    from torchsparse.utils.collate import sparse_collate_fn  # noqa: F401
    from torchsparse.utils.quantize import sparse_quantize  # noqa: F401
    assert hasattr(F, "sphash") and hasattr(F, "spdevoxelize") and hasattr(spnn, "Conv3d")
    assert hasattr(torchsparse, "cat") and hasattr(torchsparse.backend, "convolution_forward_cuda")
    for k in [k for k in sys.modules if k == "torchsparse" or k.startswith("torchsparse.")]:
        del sys.modules[k]


def test_minkunet_state_dict_matches_reference_inventory():
    """tests/golden/model_keys.json: key -> shape of the reference's MinkUNet-34 cr1.0 (make_golden_model.py);
    a reference checkpoint therefore loads with strict=True."""
    import json
    from openpcseg_b200.segmentors import MinkUNet, minkunet34_config
    with open(os.path.join(os.path.dirname(__file__), "golden", "model_keys.json")) as f:
        g = json.load(f)
    net = MinkUNet(minkunet34_config())
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert mine == g["state_dict"]
    assert sum(p.numel() for p in net.parameters()) == g["n_params"] == 37882900
