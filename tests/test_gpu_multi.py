"""2-GPU test of BASELINE config 5's shape (skipped with < 2 GPUs): rank 0 owns a batch of raw
blocks, ONE NCCL scatter of contiguous block ranges, every rank compresses + decompresses its
shard on its own GPU through the C ABI, ONE gather(v) of the compressed blocks back to rank 0,
which checks them bit-exactly against the oracle."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["SNP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from snappier_b200 import sharding
from snappier_b200.batch import Engine
from tests import helpers as H
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(rank)
N = 301
blocks = H.synthetic_blocks(777, N)
if rank == 0:
    base = torch.from_numpy(np.frombuffer(b"".join(blocks), np.uint8).copy()).to(dev)
    off = torch.arange(N, device=dev, dtype=torch.int64) * 65536
    ln = torch.full((N,), 65536, dtype=torch.int32, device=dev)
    t = (base, off, ln)
else:
    t = (None, None, None)
my_base, my_off, my_len, first, n_total = sharding.scatter_batch(*t, src=0, device=dev)
n = my_off.numel()
lo, hi = sharding.shard_range(N, world, rank)
assert (first, n_total, n) == (lo, N, hi - lo)
pitch = 76496
slots = torch.zeros(max(n, 1) * pitch, dtype=torch.uint8, device=dev)
s_off = torch.arange(n, device=dev, dtype=torch.int64) * pitch
s_cap = torch.full((n,), pitch, dtype=torch.int32, device=dev)
s_len = torch.zeros(n, dtype=torch.int32, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
eng.compress_batch_device(my_base, my_off, my_len, slots, s_off, s_cap, s_len, st, 0, stream)
out = torch.zeros(max(n, 1) * 65536, dtype=torch.uint8, device=dev)
o_len = torch.zeros(n, dtype=torch.int32, device=dev); st2 = torch.zeros(n, dtype=torch.int32, device=dev)
eng.decompress_batch_device(slots, s_off, s_len, out, my_off, my_len, o_len, st2, stream)
torch.cuda.synchronize()
assert int(st.abs().sum()) == 0 and int(st2.abs().sum()) == 0
assert torch.equal(out[: n * 65536], my_base[: n * 65536])          # round trip on this rank's shard
g_base, g_off, g_len = sharding.gather_batch(slots, s_off, s_len, dst=0, engine=eng)   # slots packed by snp_pack_batch
if rank == 0:
    from oracle import pyoracle as O
    gb = g_base.cpu().numpy(); go = g_off.cpu().numpy(); gl = g_len.cpu().numpy()
    assert len(gl) == N
    for i in range(N):
        assert gb[go[i]: go[i] + gl[i]].tobytes() == O.compress(blocks[i])[1], i   # order + bytes
    print("MULTI-GPU OK", world, "ranks", N, "blocks")
dist.destroy_process_group()
'''


@pytest.mark.gpu
@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_scatter_compress_decompress_gather_nccl(tmp_path):
    import subprocess
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SNP_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "MULTI-GPU OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
