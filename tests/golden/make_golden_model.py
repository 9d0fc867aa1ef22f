#!/usr/bin/env python
"""Freeze the parameter / buffer inventory of the reference's MinkUNet-34 cr1.0 (build container only):
state_dict key -> shape, parameter count.  The reference model is instantiated from
pcseg/model/segmentor/voxel/minkunet/minkunet.py with the MODEL block of
tools/cfgs/voxel/semantic_kitti/minkunet_mk34_cr10.yaml; modules it imports but never calls while being
constructed are stubbed.   python tests/golden/make_golden_model.py -> model_keys.json"""
import importlib
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as MG                                             # noqa: E402


class Cfg(dict):
    __getattr__ = dict.__getitem__


def import_with_stubs(name, tries=30):
    for _ in range(tries):
        try:
            return importlib.import_module(name)
        except ModuleNotFoundError as exc:
            missing = exc.name
            if missing is None or missing.startswith("pcseg") or missing.startswith("tools"):
                raise
            stub = types.ModuleType(missing)
            stub.__path__ = []
            if missing == "easydict":
                stub.EasyDict = Cfg
            sys.modules[missing] = stub
            print("stubbed", missing)
    raise RuntimeError("too many missing modules")


def main():
    import yaml
    tmp = tempfile.mkdtemp()
    MG.import_reference(tmp)
    sys.path.insert(0, "/root/reference")
    mod = import_with_stubs("pcseg.model.segmentor.voxel.minkunet.minkunet")
    with open("/root/reference/tools/cfgs/voxel/semantic_kitti/minkunet_mk34_cr10.yaml") as f:
        y = yaml.safe_load(f)
    model_cfg = Cfg(y["MODEL"])
    model_cfg["IF_DIST"] = False
    print("MODEL cfg:", dict(model_cfg))
    net = mod.MinkUNet(model_cfg, num_class=20)
    sd = net.state_dict()
    inv = {k: list(v.shape) for k, v in sd.items()}
    n_params = sum(p.numel() for p in net.parameters())
    with open(os.path.join(HERE, "model_keys.json"), "w") as f:
        json.dump({"state_dict": inv, "n_params": n_params, "cfg": {k: model_cfg[k] for k in model_cfg}}, f, indent=0)
    print(len(inv), "entries,", n_params, "parameters")


if __name__ == "__main__":
    main()
