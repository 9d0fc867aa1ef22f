"""Build libb2s.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

``python -m openpcseg_b200.build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU.  Objects are cached under ``openpcseg_b200/csrc/_obj``
keyed on the source mtime; the shared library lands at ``openpcseg_b200/libb2s.so``
(git-ignored, travels with gpurun snapshots).
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(PKG, "libb2s.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wno-deprecated-declarations",
         "-diag-suppress", "1444", "--expt-relaxed-constexpr"]


def _newer(a: str, others) -> bool:
    if not os.path.exists(a):
        return False
    t = os.path.getmtime(a)
    return all(os.path.getmtime(o) <= t for o in others)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG, "..", "include", "b2s.h")]

    def cc(src: str) -> str:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or not _newer(obj, [src] + hdrs):
            cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(cc, srcs))
    if force or not _newer(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
