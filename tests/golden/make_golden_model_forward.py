#!/usr/bin/env python
"""Run the REFERENCE MinkUNet-34 cr1.0 (pcseg/model/segmentor/voxel/minkunet/minkunet.py, on the CPU build
of its bundled torchsparse) forward on a small synthetic scan with key-seeded weights
(oracle/det_weights.py) and freeze input + per-point logits (build container only).
python tests/golden/make_golden_model_forward.py -> model_forward.npz"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as MG                                             # noqa: E402
import make_golden_model as MM                                       # noqa: E402
from oracle.det_weights import fill_                                 # noqa: E402
from openpcseg_b200.synthetic import make_scan                       # noqa: E402


def main():
    import yaml
    tmp = tempfile.mkdtemp()
    ts, _ = MG.import_reference(tmp)
    sys.path.insert(0, "/root/reference")
    mod = MM.import_with_stubs("pcseg.model.segmentor.voxel.minkunet.minkunet")
    with open("/root/reference/tools/cfgs/voxel/semantic_kitti/minkunet_mk34_cr10.yaml") as f:
        cfg = MM.Cfg(yaml.safe_load(f)["MODEL"])
    cfg["IF_DIST"] = False
    torch.manual_seed(0)
    net = mod.MinkUNet(cfg, num_class=20)
    net.load_state_dict(fill_(net.state_dict()), strict=True)
    net.train()                                                      # batch statistics, dropout p = 0
    scan = make_scan(3, n_beams=64, n_azimuth=150)
    coords = np.concatenate([scan["coords"], np.zeros((scan["coords"].shape[0], 1), np.int32)], 1)
    feats = scan["feats"]
    captured = {}
    net.classifier.register_forward_hook(lambda m, i, o: captured.__setitem__("logits", o.detach().clone()))
    lidar = ts.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords))
    try:
        net({"lidar": lidar, "targets": ts.SparseTensor(torch.from_numpy(scan["labels"]), torch.from_numpy(coords)),
             "offset": torch.tensor([coords.shape[0]]).int()})
    except Exception as exc:                                         # the loss branch moves targets to CUDA
        print("forward stopped after the classifier:", type(exc).__name__, str(exc)[:80])
    logits = captured["logits"].numpy()
    np.savez_compressed(os.path.join(HERE, "model_forward.npz"), coords=coords, feats=feats, logits=logits)
    print("voxels", coords.shape[0], "logits", logits.shape, float(np.abs(logits).max()))


if __name__ == "__main__":
    main()
