"""The batched, sync-free Lovasz-softmax + CE equals the reference-style per-class loop."""
import torch

from oracle.cpu_minkunet import CpuMinkUNet
from openpcseg_b200.segmentors.losses import SegLoss


def test_segloss_matches_reference_style_loop():
    torch.manual_seed(0)
    logits = torch.randn(20000, 20, requires_grad=True)
    target = torch.randint(0, 20, (20000,))
    target[target == 7] = 3                                   # one absent class
    mine = SegLoss(ignore_index=0, label_smoothing=0.1)(logits, target)
    ref_net = CpuMinkUNet({})
    l2 = logits.detach().clone().requires_grad_(True)
    ref = ref_net.loss(l2, target)
    assert abs(float(mine) - float(ref)) < 1e-5 * abs(float(ref))
    mine.backward()
    ref.backward()
    assert (logits.grad - l2.grad).abs().max() < 1e-6


def test_segloss_cross_entropy_term_matches_torch():
    torch.manual_seed(1)
    logits = torch.randn(5000, 20, requires_grad=True)
    target = torch.randint(0, 20, (5000,))
    for eps in (0.0, 0.1):
        mine = SegLoss(ignore_index=0, label_smoothing=eps, lovasz_weight=0.0)
        l1 = logits.detach().clone().requires_grad_(True)
        l2 = logits.detach().clone().requires_grad_(True)
        a = mine.ce_weight * mine(l1, target)
        b = torch.nn.CrossEntropyLoss(ignore_index=0, label_smoothing=eps)(l2, target)
        assert abs(float(a) - float(b)) < 1e-6 * abs(float(b)) + 1e-7
        a.backward()
        b.backward()
        assert (l1.grad - l2.grad).abs().max() < 1e-7
