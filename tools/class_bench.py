#!/usr/bin/env python3
"""Per-data-class decompress throughput of selected kernel variants (diagnostics, not the headline bench).

    python tools/class_bench.py [--blocks 32768] [--variants 7,7w4096,5] [--out gpurun_out/class_bench.json]

Classes are bench.py's 'Silesia-mix synthetic' components, one class per batch; every timed output is
verified against the raw blocks' checksums.  Also times small mixed batches (launch-latency regime).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench as B  # noqa: E402

CLASSES = [("text", 0), ("markup", 30), ("kppkn", 55), ("lz_synth", 67), ("records", 80), ("jpeg", 90), ("prng", 95),
           ("mix", None)]


def engine_with(env):
    from snappier_b200.batch import Engine
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return Engine(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def prepare(torch, engine, n, dev, force_class):
    orig = B.make_blocks
    if force_class is not None:
        B.make_blocks = lambda t, c, fb, cnt, d: orig(t, c, fb, cnt, d, force_class=force_class)
    try:
        return B.prepare_batch(torch, engine, n, 0, dev, cap_ratio=1.02 if force_class is not None and force_class >= 90 else 0.70)
    finally:
        B.make_blocks = orig


def time_decompress(torch, engine, comp, c_off, c_len, sums, weights, n, dev, reps=3):
    out = torch.empty(n * B.BLOCK, dtype=torch.uint8, device=dev)
    o_off = torch.arange(n, device=dev, dtype=torch.int64) * B.BLOCK
    o_cap = torch.full((n,), B.BLOCK, dtype=torch.int32, device=dev)
    o_len = torch.zeros(n, dtype=torch.int32, device=dev)
    status = torch.full((n,), -9, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        engine.decompress_batch_device(comp, c_off[:n], c_len[:n], out, o_off, o_cap, o_len, status, stream)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        out.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.decompress_batch_device(comp, c_off[:n], c_len[:n], out, o_off, o_cap, o_len, status, stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ok = int(status.abs().sum()) == 0 and bool((o_len == B.BLOCK).all()) and \
        torch.equal(B.block_checksums(torch, out, weights), sums[:n])
    return best, ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=1 << 15)
    ap.add_argument("--variants", default="7,7w2048,5")
    ap.add_argument("--small", default="64,256,1024,4096")
    ap.add_argument("--classes", default="", help="comma-separated subset of class names (default: all)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "class_bench.json"))
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    variants = args.variants.split(",")
    def env_of(v):  # "7" = default kernel, "7w4096" = its 4 KiB-window instantiation, "5" = the round-1 kernel
        k, _, w = v.partition("w")
        w, _, c = w.partition("c")
        if k.startswith("8"):  # "8" = lane-per-block engine, "8c1" / "8c2" = its other instantiations (SNP_V8_CFG)
            return {"SNP_DECOMP_KERNEL": "8", "SNP_V8_CFG": c or "0"}
        return {"SNP_DECOMP_KERNEL": k, "SNP_V7_WINDOW": w or "4096", "SNP_V7_CTAS": c or "0"}
    engines = {v: engine_with(env_of(v)) for v in variants}
    prep_engine = engines[variants[0]]
    res = {"blocks": args.blocks, "classes": {}, "small_mix": {}}
    wanted = [c for c in args.classes.split(",") if c]
    for name, fc in CLASSES:
        if wanted and name not in wanted:
            continue
        comp, c_off, c_len, sums, weights, cbytes = prepare(torch, prep_engine, args.blocks, dev, fc)
        row = {"ratio": round(cbytes / (args.blocks * B.BLOCK), 4)}
        for v, e in engines.items():
            ms, ok = time_decompress(torch, e, comp, c_off, c_len, sums, weights, args.blocks, dev)
            row[f"v{v}_GBps"] = round(args.blocks * B.BLOCK / ms / 1e6, 1)
            row[f"v{v}_ok"] = ok
        res["classes"][name] = row
        print(name, row, flush=True)
        if name == "mix":
            for ns in [int(x) for x in args.small.split(",") if x]:
                ns = min(ns, args.blocks)
                r = {}
                for v, e in engines.items():
                    ms, ok = time_decompress(torch, e, comp, c_off, c_len, sums, weights, ns, dev, reps=5)
                    r[f"v{v}_ms"] = round(ms, 4)
                    r[f"v{v}_ok"] = ok
                res["small_mix"][str(ns)] = r
                print("small", ns, r, flush=True)
        del comp
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    for e in engines.values():
        e.close()


if __name__ == "__main__":
    main()
