"""Builds snappier_b200/libsnappier_b200.so (the C ABI of include/snappier_b200.h)
with nvcc for sm_100a.  In-tree on purpose: the .so travels to the GPU box with
the repo snapshot.  `python -m snappier_b200.build [--force]`."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libsnappier_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "-cudart", "static",
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def deps() -> list[str]:
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
        os.path.join(os.path.dirname(PKG), "include", "snappier_b200.h"), os.path.abspath(__file__)]


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    cmd = [NVCC] + FLAGS + ["-o", SO] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(PKG, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libsnappier_b200.so")
    if verbose:
        print(log)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
