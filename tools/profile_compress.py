#!/usr/bin/env python3
"""Runs a few compress launches of one data class (for ncu): python tools/profile_compress.py <text|mix|config3> <blocks>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import bench as B  # noqa: E402
import class_bench as CB  # noqa: E402

name, n = sys.argv[1], int(sys.argv[2])
dev = torch.device("cuda", 0)
eng = CB.engine_with({})
corpus_dev = {k: torch.from_numpy(v).to(dev) for k, v in B.load_corpus().items()}
if name == "config3":
    raw = torch.cat([B.make_blocks_config3(torch, b0, min(8192, n - b0), dev) for b0 in range(0, n, 8192)])
else:
    fc = {"text": 0, "mix": None}[name]
    raw = torch.cat([B.make_blocks(torch, corpus_dev, b0, min(8192, n - b0), dev, force_class=fc) for b0 in range(0, n, 8192)])
slots = torch.empty(n * B.PITCH, dtype=torch.uint8, device=dev)
idx = torch.arange(n, device=dev, dtype=torch.int64)
r_len = torch.full((n,), B.BLOCK, dtype=torch.int32, device=dev)
s_cap = torch.full((n,), B.PITCH, dtype=torch.int32, device=dev)
s_len = torch.zeros(n, dtype=torch.int32, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.compress_batch_device(raw.view(-1), idx * B.BLOCK, r_len, slots, idx * B.PITCH, s_cap, s_len, st, 0,
                              torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
print(name, n, "ms", e0.elapsed_time(e1), "GB/s", n * B.BLOCK / e0.elapsed_time(e1) / 1e6, "ok", int(st.abs().sum()) == 0)
