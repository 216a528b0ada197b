#!/usr/bin/env python3
"""e2e decompress (pinned host buffers) for several host-pipeline chunk sizes: python tools/e2e_probe.py [blocks]"""
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
dev = torch.device("cuda", 0)
eng0 = CB.engine_with({})
comp, c_off, c_len, sums, weights, cbytes = B.prepare_batch(torch, eng0, n, 0, dev)
h_in = torch.empty(cbytes, dtype=torch.uint8).pin_memory()
h_in.copy_(comp[:cbytes])
h_out = torch.empty(n * B.BLOCK, dtype=torch.uint8).pin_memory()
off = c_off.cpu().numpy().astype(np.uint64)
ln = c_len.cpu().numpy().astype(np.uint32)
ooff = np.arange(n, dtype=np.uint64) * B.BLOCK
ocap = np.full(n, B.BLOCK, np.uint32)
for mb, early in ((128, 1), (64, 1), (256, 1)):
    eng = CB.engine_with({"SNP_HOST_CHUNK_MB": str(mb), "SNP_HOST_EARLY_D2H": str(early)})
    for _ in range(2):
        eng.decompress_batch_host(h_in.numpy(), off, ln, h_out.numpy(), ooff, ocap)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ol, st = eng.decompress_batch_host(h_in.numpy(), off, ln, h_out.numpy(), ooff, ocap)
    dt = (time.perf_counter() - t0) / 5
    good = not st.any() and torch.equal(B.block_checksums(torch, h_out.to(dev), weights), sums[:n])
    print(f"chunk {mb:4d} MiB early_d2h={early}: {n * B.BLOCK / dt / 1e9:6.2f} GB/s  ok={good}", flush=True)
    eng.close()
