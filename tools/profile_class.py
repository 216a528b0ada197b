#!/usr/bin/env python3
"""Runs one decompress of a single-class batch (for ncu): python tools/profile_class.py <class> <blocks> [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import bench as B  # noqa: E402
import class_bench as CB  # noqa: E402

name, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
eng = CB.engine_with({})
fc = dict(CB.CLASSES)[name]
comp, c_off, c_len, sums, weights, cbytes = CB.prepare(torch, eng, n, dev, fc)
ms, ok = CB.time_decompress(torch, eng, comp, c_off, c_len, sums, weights, n, dev, reps=reps)
print(name, n, "ms", ms, "GB/s", n * B.BLOCK / ms / 1e6, "ok", ok, "ratio", cbytes / (n * B.BLOCK))
