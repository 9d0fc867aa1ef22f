#!/usr/bin/env python
"""Freeze the reference's training loss on seeded logits (build container only): CrossEntropyLoss(ignore_index,
label_smoothing) + tools/utils/common/lovasz_losses.lovasz_softmax(softmax, target, ignore) as combined in
pcseg/loss/__init__.py:113-122 with weights [1, 1].   python tests/golden/make_golden_loss.py -> loss.npz"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_lovasz", "/root/reference/tools/utils/common/lovasz_losses.py")
    lov = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lov)
    out = {}
    torch.manual_seed(0)
    for tag, n, smoothing in (("a", 4000, 0.0), ("b", 6000, 0.1)):
        logits = (torch.randn(n, 20) * 2).requires_grad_(True)
        target = torch.randint(0, 20, (n,))
        target[target == 11] = 4                                    # one absent class
        ce = torch.nn.CrossEntropyLoss(ignore_index=0, label_smoothing=smoothing)(logits, target)
        lv = lov.lovasz_softmax(logits.softmax(dim=1), target, ignore=0)
        loss = ce * 1.0 + lv * 1.0
        loss.backward()
        out.update({f"{tag}_logits": logits.detach().numpy(), f"{tag}_target": target.numpy(),
                    f"{tag}_ce": ce.detach().numpy(), f"{tag}_lovasz": lv.detach().numpy(),
                    f"{tag}_loss": loss.detach().numpy(), f"{tag}_grad": logits.grad.numpy(),
                    f"{tag}_smoothing": np.array(smoothing)})
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
