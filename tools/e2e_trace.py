import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import numpy as np, torch
import bench as B, class_bench as CB
n = 32768
dev = torch.device("cuda", 0)
eng0 = CB.engine_with({})
comp, c_off, c_len, sums, weights, cbytes = B.prepare_batch(torch, eng0, n, 0, dev)
h_in = torch.empty(cbytes, dtype=torch.uint8).pin_memory(); h_in.copy_(comp[:cbytes])
h_out = torch.empty(n * B.BLOCK, dtype=torch.uint8).pin_memory()
off = c_off.cpu().numpy().astype(np.uint64); ln = c_len.cpu().numpy().astype(np.uint32)
ooff = np.arange(n, dtype=np.uint64) * B.BLOCK; ocap = np.full(n, B.BLOCK, np.uint32)
eng = CB.engine_with({})
for _ in range(2): eng.decompress_batch_host(h_in.numpy(), off, ln, h_out.numpy(), ooff, ocap)
engt = CB.engine_with({"SNP_HOST_TRACE": "1"})
for _ in range(2): engt.decompress_batch_host(h_in.numpy(), off, ln, h_out.numpy(), ooff, ocap)
