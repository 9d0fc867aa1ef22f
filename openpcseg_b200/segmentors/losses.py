"""Segmentation loss of the voxel segmentors: cross-entropy + Lovasz-softmax
(pcseg/loss/__init__.py:15-137 defaults, tools/utils/common/lovasz_losses.py:158-205).

The Lovasz term is computed for all classes in one batched sort and without any host
synchronisation: ignored points get zero error and zero weight instead of being filtered
out (same value: a zero-error, zero-weight entry changes no cumulative sum), and the
'present classes only' average is a masked mean.
"""
import torch
from torch import nn
import torch.nn.functional as TF

__all__ = ["SegLoss", "lovasz_softmax_flat"]


def _cumsum_long_rows(x: torch.Tensor, block: int = 2048) -> torch.Tensor:
    """Inclusive cumsum along dim 1 of a [C, P] tensor with few, very long rows.  torch.cumsum runs
    one thread block per row there (1.3 ms for [20, 760k] on B200 = 4.6 % of the training step); a
    two-level scan (inside 2048-element blocks, then over the block totals) keeps the GPU busy."""
    c, p = x.shape
    if p <= 4 * block:
        return x.cumsum(1)
    pad = (-p) % block
    xp = torch.nn.functional.pad(x, (0, pad)).view(c, -1, block)
    inner = xp.cumsum(2)
    offsets = inner[:, :, -1].cumsum(1) - inner[:, :, -1]
    return (inner + offsets[:, :, None]).view(c, -1)[:, :p]


def lovasz_softmax_flat(probs: torch.Tensor, labels: torch.Tensor, ignore_index: int) -> torch.Tensor:
    """probs [P, C], labels [P].  Works class-major ([C, P], scans along the contiguous axis:
    torch's outer-dimension cumsum is ~50x slower on a [190k, 20] tensor)."""
    n_cls = probs.shape[1]
    valid = (labels != ignore_index).to(probs.dtype)[None, :]                       # [1, P]
    cls = torch.arange(n_cls, device=probs.device)[:, None]
    fg = (labels[None, :] == cls).to(probs.dtype) * valid                            # [C, P]
    errors = (fg - probs.t()).abs() * valid
    errors_sorted, perm = torch.sort(errors, dim=1, descending=True)
    fg_sorted = torch.gather(fg, 1, perm)
    bg_sorted = torch.gather((1.0 - fg) * valid, 1, perm)
    gts = fg_sorted.sum(1, keepdim=True)
    inter = gts - _cumsum_long_rows(fg_sorted)
    union = gts + _cumsum_long_rows(bg_sorted)
    jaccard = 1.0 - inter / union.clamp_min(1e-12)
    grad = torch.cat([jaccard[:, :1], jaccard[:, 1:] - jaccard[:, :-1]], 1)
    per_class = (errors_sorted * grad).sum(1)
    present = (gts.squeeze(1) > 0).to(probs.dtype)
    return (per_class * present).sum() / present.sum().clamp_min(1.0)


class SegLoss(nn.Module):
    def __init__(self, ignore_index: int = 0, label_smoothing: float = 0.0, ce_weight: float = 1.0,
                 lovasz_weight: float = 1.0):
        super().__init__()
        self.ignore_index = ignore_index
        self.ce = nn.CrossEntropyLoss(ignore_index=ignore_index, label_smoothing=label_smoothing)
        self.ce_weight, self.lovasz_weight = ce_weight, lovasz_weight

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        logits = logits.float()
        # same value as nn.CrossEntropyLoss(ignore_index, label_smoothing) (mean reduction) built
        # from element-parallel ops: the library's nll_loss reduction is a single-CTA kernel
        # (1.6 ms forward for 1.9 M points); the log-softmax is shared with the Lovasz term
        eps = float(self.ce.label_smoothing)
        logp = TF.log_softmax(logits, dim=1)
        valid = target != self.ignore_index
        n_valid = valid.sum()
        picked = logp.gather(1, target.clamp(min=0, max=logits.shape[1] - 1).unsqueeze(1)).squeeze(1)
        ce = -(picked * valid).sum() / n_valid
        if eps > 0.0:
            smooth = -(logp.sum(dim=1) * valid).sum() / n_valid
            ce = (1.0 - eps) * ce + (eps / logits.shape[1]) * smooth
        probs = logp.exp()
        loss = self.ce_weight * ce
        loss = loss + self.lovasz_weight * lovasz_softmax_flat(probs, target, self.ignore_index)
        return loss
