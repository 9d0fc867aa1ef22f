#!/usr/bin/env python
"""Build-container check (needs /root/reference): the reference's four sparse segmentors are instantiated
UNMODIFIED on top of ``openpcseg_b200.install_as_torchsparse()`` with the MODEL block of their yaml configs.
Modules they import but never call while being constructed (SharedArray, torch_scatter, easydict, imp,
RPVNet's range_utils extension) are stubbed.  Construction needs no GPU; forward passes do.

Last run (round 1):  MinkUNet 37 882 900 params / 63 spnn.Conv3d, SPVCNN 21 785 780 / 49,
Cylinder_TS 55 892 042 / 48, RPVNet 123 021 828 / 63.
"""
import importlib
import os
import sys
import types

import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openpcseg_b200                                                # noqa: E402

openpcseg_b200.install_as_torchsparse()
sys.path.insert(0, "/root/reference")


class Cfg(dict):
    __getattr__ = dict.__getitem__


def wrap(d):
    return Cfg({k: (wrap(v) if isinstance(v, dict) else v) for k, v in d.items()})


for name in ("imp", "SharedArray", "torch_scatter", "easydict"):
    try:
        importlib.import_module(name)
    except Exception:
        stub = types.ModuleType(name)
        stub.__path__ = []
        stub.EasyDict = Cfg
        sys.modules[name] = stub
ru, run, runf = types.ModuleType("range_utils"), types.ModuleType("range_utils.nn"), \
    types.ModuleType("range_utils.nn.functional")
ru.__path__, run.__path__ = [], []
for fn in ("denselize", "map_count", "range_to_point", "point_to_range"):
    setattr(runf, fn, lambda *a, **k: None)
ru.nn, run.functional = run, runf
sys.modules.update({"range_utils": ru, "range_utils.nn": run, "range_utils.nn.functional": runf})

import torchsparse.nn as spnn                                        # noqa: E402  (the shim)

CASES = [("voxel/semantic_kitti/minkunet_mk34_cr10.yaml", "pcseg.model.segmentor.voxel.minkunet.minkunet", "MinkUNet"),
         ("fusion/semantic_kitti/spvcnn_mk18_cr10.yaml", "pcseg.model.segmentor.fusion.spvcnn.spvcnn", "SPVCNN"),
         ("voxel/semantic_kitti/cylinder_cy480_cr10.yaml", "pcseg.model.segmentor.voxel.cylinder3d.cylinder_ts",
          "Cylinder_TS"),
         ("fusion/semantic_kitti/rpvnet_mk34_cr17_5.yaml", "pcseg.model.segmentor.fusion.rpvnet.rpvnet", "RPVNet")]
for cfg_file, module, cls in CASES:
    with open("/root/reference/tools/cfgs/" + cfg_file) as f:
        cfg = wrap(yaml.safe_load(f)["MODEL"])
    cfg["IF_DIST"] = False
    net = getattr(importlib.import_module(module), cls)(cfg, 20)
    n_conv = sum(isinstance(m, spnn.Conv3d) for m in net.modules())
    assert type(net).__module__ == module and spnn.__name__.startswith("openpcseg_b200")
    print(f"{cls:12s} constructed on the shim: {sum(p.numel() for p in net.parameters()):>11,d} parameters, "
          f"{n_conv} spnn.Conv3d")
