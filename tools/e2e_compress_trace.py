#!/usr/bin/env python3
"""Per-chunk timeline of the host-mode compress pipeline (SNP_HOST_TRACE): python tools/e2e_compress_trace.py [blocks]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench as B  # noqa: E402
import class_bench as CB  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
raw = torch.cat([B.make_blocks_config3(torch, b0, min(8192, n - b0), dev) for b0 in range(0, n, 8192)]).view(-1)
h_raw = torch.empty(n * B.BLOCK, dtype=torch.uint8).pin_memory()
h_raw.copy_(raw)
h_slots = torch.empty(n * B.PITCH, dtype=torch.uint8).pin_memory()
r_off = np.arange(n, dtype=np.uint64) * B.BLOCK
r_len = np.full(n, B.BLOCK, np.uint32)
s_off = np.arange(n, dtype=np.uint64) * B.PITCH
s_cap = np.full(n, B.PITCH, np.uint32)
eng = CB.engine_with({"SNP_HOST_TRACE": "0"})
for _ in range(2):
    eng.compress_batch_host(h_raw.numpy(), r_off, r_len, h_slots.numpy(), s_off, s_cap, 0)
eng.close()
eng = CB.engine_with({"SNP_HOST_TRACE": "1"})
eng.compress_batch_host(h_raw.numpy(), r_off, r_len, h_slots.numpy(), s_off, s_cap, 0)
t0 = time.perf_counter()
eng.compress_batch_host(h_raw.numpy(), r_off, r_len, h_slots.numpy(), s_off, s_cap, 0)
print("call", time.perf_counter() - t0, "s", n * B.BLOCK / (time.perf_counter() - t0) / 1e9, "GB/s")
