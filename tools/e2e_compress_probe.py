#!/usr/bin/env python3
"""e2e compress (pinned host buffers, config-3 blocks) for several host-pipeline chunk sizes:
python tools/e2e_compress_probe.py [blocks]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
kind = sys.argv[2] if len(sys.argv) > 2 else "config3"
dev = torch.device("cuda", 0)
if kind == "config3":
    raw = torch.cat([B.make_blocks_config3(torch, b0, min(8192, n - b0), dev) for b0 in range(0, n, 8192)]).view(-1)
else:
    corpus_dev = {k: torch.from_numpy(v).to(dev) for k, v in B.load_corpus().items()}
    raw = torch.cat([B.make_blocks(torch, corpus_dev, b0, min(8192, n - b0), dev) for b0 in range(0, n, 8192)]).view(-1)
h_raw = torch.empty(n * B.BLOCK, dtype=torch.uint8).pin_memory()
h_raw.copy_(raw)
h_slots = torch.empty(n * B.PITCH, dtype=torch.uint8).pin_memory()
r_off = np.arange(n, dtype=np.uint64) * B.BLOCK
r_len = np.full(n, B.BLOCK, np.uint32)
s_off = np.arange(n, dtype=np.uint64) * B.PITCH
s_cap = np.full(n, B.PITCH, np.uint32)
ref = None
for mb in (64, 128, 256, 512):
    eng = CB.engine_with({"SNP_HOST_COMP_CHUNK_MB": str(mb)})
    for _ in range(2):
        eng.compress_batch_host(h_raw.numpy(), r_off, r_len, h_slots.numpy(), s_off, s_cap, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        ol, st = eng.compress_batch_host(h_raw.numpy(), r_off, r_len, h_slots.numpy(), s_off, s_cap, 0)
    dt = (time.perf_counter() - t0) / 4
    tot = int(ol.astype(np.int64).sum())
    ref = ref or tot
    print(f"{kind} chunk {mb:4d} MiB: {n * B.BLOCK / dt / 1e9:6.2f} GB/s  ok={not st.any() and tot == ref}", flush=True)
    eng.close()
