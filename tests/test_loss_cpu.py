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


def test_segloss_matches_reference_loss_outputs():
    """tests/golden/loss.npz: CE + Lovasz of the reference's own functions (make_golden_loss.py)."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss.npz"))
    for tag in ("a", "b"):
        logits = torch.from_numpy(g[f"{tag}_logits"]).requires_grad_(True)
        target = torch.from_numpy(g[f"{tag}_target"])
        crit = SegLoss(ignore_index=0, label_smoothing=float(g[f"{tag}_smoothing"]))
        loss = crit(logits, target)
        assert abs(float(loss) - float(g[f"{tag}_loss"])) < 2e-6 * abs(float(g[f"{tag}_loss"]))
        loss.backward()
        assert float((logits.grad - torch.from_numpy(g[f"{tag}_grad"])).abs().max()) < 1e-7
        only_lov = SegLoss(ignore_index=0, label_smoothing=0.0, ce_weight=0.0)(logits.detach(), target)
        assert abs(float(only_lov) - float(g[f"{tag}_lovasz"])) < 2e-6 * abs(float(g[f"{tag}_lovasz"]))
