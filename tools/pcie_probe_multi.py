#!/usr/bin/env python3
"""Concurrent PCIe bandwidth of all ranks of one box (run under torchrun): H2D alone, D2H alone and both at once from
pinned host buffers, every rank on its own GPU at the same time, with and without binding the rank to its GPU's NUMA node.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe_multi.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from snappier_b200 import numa  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 1 << 30
d_a = torch.empty(N, dtype=torch.uint8, device=dev)
d_b = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def measure(tag):
    h_a = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_a.fill_(1)
    h_b.fill_(2)
    res = {}
    for mode in ("h2d", "d2h", "duplex"):
        for it in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if mode in ("h2d", "duplex"):
                with torch.cuda.stream(s1):
                    s1.wait_event(e0)
                    for _ in range(4):
                        d_a.copy_(h_a, non_blocking=True)
            if mode in ("d2h", "duplex"):
                with torch.cuda.stream(s2):
                    s2.wait_event(e0)
                    for _ in range(4):
                        h_b.copy_(d_b, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_dir = 4 * N / (float(t[0]) * 1e-3) / 1e9
        res[mode] = round(per_dir * world, 1)  # aggregate GB/s per direction (max-over-ranks time)
    if rank == 0:
        print(json.dumps({"probe": tag, "world": world, "aggregate_GBps_per_direction": res}), flush=True)
    del h_a, h_b


measure("unbound")
info = numa.bind_to_gpu_node(local)
allinfo = [None] * world
dist.all_gather_object(allinfo, info)
if rank == 0:
    print(json.dumps({"numa": allinfo}), flush=True)
measure("bound to the GPU's NUMA node")
dist.destroy_process_group()
