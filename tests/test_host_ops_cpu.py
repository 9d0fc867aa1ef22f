"""Host-side tensor ops that need no CUDA: channel concat with the vectorised backward."""
import pytest
import torch

from openpcseg_b200.torchsparse.operators import _CatFeats, _slice_cols


@pytest.mark.parametrize("dtype,widths", [(torch.float16, (96, 32)), (torch.float16, (256, 128, 96)),
                                          (torch.float32, (5, 3)), (torch.float16, (8, 4)), (torch.float32, (4, 4, 4))])
def test_cat_feats_forward_backward(dtype, widths):
    torch.manual_seed(0)
    parts = [torch.randn(37, w).to(dtype).requires_grad_(True) for w in widths]
    y = _CatFeats.apply(*parts)
    assert torch.equal(y, torch.cat([p.detach() for p in parts], 1))
    g = torch.randn(37, sum(widths)).to(dtype)
    y.backward(g)
    c0 = 0
    for p, w in zip(parts, widths):
        assert p.grad.is_contiguous() and torch.equal(p.grad, g[:, c0:c0 + w])
        c0 += w


def test_slice_cols_paths_and_partial_grads():
    g = torch.arange(40 * 24, dtype=torch.float32).view(40, 24).half()
    assert torch.equal(_slice_cols(g, 8, 24), g[:, 8:24])            # 16-byte aligned window: vector view
    assert torch.equal(_slice_cols(g, 3, 7), g[:, 3:7])              # unaligned: plain narrow
    assert torch.equal(_slice_cols(g[:, :20], 0, 8), g[:, :8])       # non-contiguous input
    a = torch.randn(5, 8, requires_grad=True)
    b = torch.randn(5, 8)                                            # no grad wanted for b
    _CatFeats.apply(a, b).sum().backward()
    assert torch.equal(a.grad, torch.ones(5, 8))


def test_weight_bank_host_logic(monkeypatch):
    """functional._WeightBank without a GPU (the launch is recorded, not run): one single-entry launch at first
    sight, ONE launch over every live parameter when a stale one is looked up, descriptor prefix sums in 32 x 32
    tiles, dead parameters leave the table, nothing is attached to the parameter (it stays picklable)."""
    import gc
    import pickle
    import openpcseg_b200.torchsparse.nn.functional as F
    calls = []
    monkeypatch.setattr(F.B, "weights_refresh", lambda table, total: calls.append((table.clone(), total)))
    bank = F._WeightBank(torch.device("cpu"))
    ps = [torch.nn.Parameter(torch.randn(27, 32, 64)), torch.nn.Parameter(torch.randn(8, 40, 72)),
          torch.nn.Parameter(torch.randn(64, 128))]
    es = [bank.lookup(p) for p in ps]
    assert [(c[0].shape[0], c[1]) for c in calls] == [(1, 27 * 1 * 2), (1, 8 * 2 * 3), (1, 1 * 2 * 4)]
    assert es[2].kmajor.shape == (1, 128, 64) and es[1].cast.shape == (8, 40, 72)
    assert bank.lookup(ps[0]) is es[0] and len(calls) == 3                 # fresh: no launch
    with torch.no_grad():
        ps[1].add_(1.0)
    assert bank.lookup(ps[1]) is es[1]
    table, total = calls[-1]
    assert table.shape == (3, 5) and total == 54 + 48 + 8
    assert table[:, 0].tolist() == [p.data_ptr() for p in ps]
    assert (table[:, 3] & 0xFFFFFFFF).tolist() == [27, 8, 1] and (table[:, 3] >> 32).tolist() == [32, 40, 64]
    assert (table[:, 4] & 0xFFFFFFFF).tolist() == [64, 72, 128] and (table[:, 4] >> 32).tolist() == [0, 54, 102]
    n = len(calls)
    bank.lookup(ps[2])
    assert len(calls) == n                                                  # refreshed together with ps[1]
    del es
    gone = ps.pop(1)
    del gone
    gc.collect()
    with torch.no_grad():
        ps[0].mul_(2.0)
    bank.lookup(ps[0])
    assert calls[-1][0].shape[0] == 2 and calls[-1][1] == 54 + 8 and len(bank.by_id) == 2
    pickle.dumps(ps[0])
    assert not F._bankable(torch.randn(27, 32, 32), torch.float16)         # not a Parameter / not CUDA


def test_zero_sums_and_cached_offsets_host_logic():
    import openpcseg_b200.torchsparse.nn.functional as F
    from openpcseg_b200.torchsparse.nn.utils import get_kernel_offsets, kernel_offsets_cached
    seen = [F.zero_sums(32 * (1 + i % 8), "cpu") for i in range(600)]      # several chunks
    for i, s in enumerate(seen):
        assert s.shape == (2, 32 * (1 + i % 8)) and s.dtype == torch.float64 and s.is_contiguous()
        assert float(s.abs().sum()) == 0.0
        s.fill_(i + 1.0)
    assert all(bool((s == i + 1.0).all()) for i, s in enumerate(seen))      # slices never overlap
    big = F.zero_sums(F._ZERO_CHUNK, "cpu")                                  # wider than a chunk
    assert big.shape == (2, F._ZERO_CHUNK) and float(big.abs().sum()) == 0.0
    a = kernel_offsets_cached(3, 2, 1, "cpu")
    assert a is kernel_offsets_cached((3, 3, 3), (2, 2, 2), (1, 1, 1), torch.device("cpu"))
    assert torch.equal(a, get_kernel_offsets(3, 2, 1, "cpu")) and a is not get_kernel_offsets(3, 2, 1, "cpu")


def test_batch_norm_act_fallback_with_sparse_wrappers():
    """the fallback of batch_norm_act must run the DENSE batch norm of the module's class on
    the [N, C] rows - MinkUNet's norms are SparseTensor wrappers (segmentors/minkunet.py _SparseBN) whose own
    forward expects a SparseTensor.  Eval mode and CPU rows both take the fallback."""
    from openpcseg_b200.segmentors.minkunet import _SparseBN
    from openpcseg_b200.torchsparse.nn.functional import batch_norm_act
    torch.manual_seed(0)
    x, res = torch.randn(10, 8), torch.randn(10, 8)
    bn = _SparseBN(8)
    with torch.no_grad():
        bn.running_mean.normal_()
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.normal_()
        bn.bias.normal_()
    ref = torch.nn.BatchNorm1d(8)
    ref.load_state_dict(bn.state_dict())
    y = batch_norm_act(x, bn.eval(), relu=True, residual=res)
    assert torch.allclose(y, torch.relu(ref.eval()(x) + res), atol=1e-6)
    bn.train(), ref.train()
    y = batch_norm_act(x, bn, relu=False)                                   # CPU rows: stock training-mode path
    assert torch.allclose(y, ref(x), atol=1e-6)
    assert torch.allclose(bn.running_mean, ref.running_mean) and int(bn.num_batches_tracked) == 1
