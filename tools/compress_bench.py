#!/usr/bin/env python3
"""Compress throughput of engine settings (diagnostics): python tools/compress_bench.py [--blocks 32768]
Variants: "N" = SNP_COMP_CTAS_PER_SM, "kN" = SNP_COMP_KERNEL, "wN" = SNP_COMP_FIRST_WIDTH."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import bench as B  # noqa: E402
import class_bench as CB  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=1 << 15)
    ap.add_argument("--variants", default="8,4,2,w32,w8")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "compress_bench.json"))
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    n = args.blocks
    corpus_dev = {k: torch.from_numpy(v).to(dev) for k, v in B.load_corpus().items()}
    data = {
        "config3": torch.cat([B.make_blocks_config3(torch, b0, min(8192, n - b0), dev) for b0 in range(0, n, 8192)]),
        "text": torch.cat([B.make_blocks(torch, corpus_dev, b0, min(8192, n - b0), dev, force_class=0) for b0 in range(0, n, 8192)]),
        "mix": torch.cat([B.make_blocks(torch, corpus_dev, b0, min(8192, n - b0), dev) for b0 in range(0, n, 8192)]),
    }
    slots = torch.empty(n * B.PITCH, dtype=torch.uint8, device=dev)
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    r_off, s_off = idx * B.BLOCK, idx * B.PITCH
    r_len = torch.full((n,), B.BLOCK, dtype=torch.int32, device=dev)
    s_cap = torch.full((n,), B.PITCH, dtype=torch.int32, device=dev)
    s_len = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    res = {}
    ref_len = {}
    for v in args.variants.split(","):
        env = {"SNP_COMP_CTAS_PER_SM": "8", "SNP_COMP_KERNEL": "3"}
        if v.startswith("k"):  # "k2" = SNP_COMP_KERNEL=2, "k6c4" = kernel 6 at 4 CTAs per SM
            k, _, cc = v[1:].partition("c")
            env["SNP_COMP_KERNEL"] = k
            if cc:
                env["SNP_COMP_CTAS_PER_SM"] = cc
        elif v.startswith("w"):  # "w8" = SNP_COMP_FIRST_WIDTH=8
            env["SNP_COMP_FIRST_WIDTH"] = v[1:]
        else:
            env["SNP_COMP_CTAS_PER_SM"] = v
        eng = CB.engine_with(env)
        row = {}
        for name, raw in data.items():
            best = 1e30
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.compress_batch_device(raw.view(-1), r_off, r_len, slots, s_off, s_cap, s_len, st, 0, stream)
                e1.record()
                torch.cuda.synchronize()
                if rep:
                    best = min(best, e0.elapsed_time(e1))
            ok = int(st.abs().sum()) == 0
            tot = int(s_len.to(torch.int64).sum())
            ok = ok and ref_len.setdefault(name, tot) == tot
            row[name] = round(n * B.BLOCK / best / 1e6, 1)
            row[name + "_ok"] = ok
        res[v] = row
        print(v, row, flush=True)
        eng.close()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
