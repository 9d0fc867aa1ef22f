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
