#!/usr/bin/env python3
"""Decompress throughput on batches of SMALL blocks (the 'Silesia-mix synthetic' data of bench.py cut into S-byte
blocks, each compressed on its own): warp-per-block engine (7) against the lane-per-block engine (8).

    python tools/small_block_bench.py [--sizes 512,1024,4096,16384] [--mib 2048] [--variants 8,7]
"""
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
    ap.add_argument("--sizes", default="512,1024,4096,16384")
    ap.add_argument("--mib", type=int, default=2048)
    ap.add_argument("--variants", default="8,7")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "small_block_bench.json"))
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    variants = args.variants.split(",")

    def env_of(v):
        k, _, c = v.partition("c")
        return {"SNP_DECOMP_KERNEL": k, "SNP_V8_CFG": c or "0"}
    engines = {v: CB.engine_with(env_of(v)) for v in variants}
    eng = engines[variants[0]]
    stream = torch.cuda.current_stream().cuda_stream
    corpus_dev = {k: torch.from_numpy(v).to(dev) for k, v in B.load_corpus().items()}
    nbig = args.mib * (1 << 20) // B.BLOCK
    raw = torch.cat([B.make_blocks(torch, corpus_dev, b0, min(8192, nbig - b0), dev) for b0 in range(0, nbig, 8192)]).view(-1)
    res = {}
    for S in [int(x) for x in args.sizes.split(",")]:
        n = raw.numel() // S
        pitch = 32 + S + S // 6
        r_off = torch.arange(n, device=dev, dtype=torch.int64) * S
        r_len = torch.full((n,), S, dtype=torch.int32, device=dev)
        slots = torch.empty(n * pitch, dtype=torch.uint8, device=dev)
        s_off = torch.arange(n, device=dev, dtype=torch.int64) * pitch
        s_cap = torch.full((n,), pitch, dtype=torch.int32, device=dev)
        s_len = torch.zeros(n, dtype=torch.int32, device=dev)
        s_st = torch.zeros(n, dtype=torch.int32, device=dev)
        eng.compress_batch_device(raw, r_off, r_len, slots, s_off, s_cap, s_len, s_st, 0, stream)
        torch.cuda.synchronize()
        assert int(s_st.abs().sum()) == 0
        cbytes = int(s_len.to(torch.int64).sum())
        out = torch.empty_like(raw)
        o_cap = r_len.clone()
        o_len = torch.zeros(n, dtype=torch.int32, device=dev)
        st = torch.full((n,), -9, dtype=torch.int32, device=dev)
        row = {"blocks": n, "ratio": round(cbytes / raw.numel(), 4)}
        for v, e in engines.items():
            best = 1e30
            for rep in range(4):
                out.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                e.decompress_batch_device(slots, s_off, s_len, out, r_off, o_cap, o_len, st, stream)
                e1.record()
                torch.cuda.synchronize()
                if rep:
                    best = min(best, e0.elapsed_time(e1))
            ok = int(st.abs().sum()) == 0 and bool((o_len == S).all()) and torch.equal(out, raw)
            row[f"v{v}_GBps"] = round(raw.numel() / best / 1e6, 1)
            row[f"v{v}_ok"] = ok
        res[str(S)] = row
        print(S, row, flush=True)
        del slots, out
        torch.cuda.empty_cache()
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
