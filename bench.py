#!/usr/bin/env python3
"""bench.py -- batched 64 KiB Snappy block decompress on B200 (BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 5 --warmup 3            # our CUDA engine
    python bench.py --impl reference --steps 3 --warmup 1    # the reference algorithm on host cores

One "step" = one pass of the hot path over one batch: every rank decompresses its own
2^20 precompressed 64 KiB blocks ("Silesia-mix synthetic", see make_blocks) with ONE
kernel launch through the C ABI (snp_decompress_batch, device pointers).  `value` is
uncompressed GB/s with inputs resident in HBM; `e2e` is the same metric through the
C ABI with HOST buffers (pinned), H2D + kernel + D2H inside the timed region.
The compressed inputs are produced by our own GPU compressor (CRC32C hash mode),
which the parity tests pin to the oracle; the decompressed output of the timed runs is
verified against per-block checksums of the raw data afterwards.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 65536
PITCH = 76496  # Snappy.GetMaxCompressedLength(65536)
METRIC = "uncompressed GB/s (batched decompress, 64 KiB blocks)"
DECOMPRESS_KERNEL = "snp::k_decompress_v7<4096, 8, 4>"   # the launches behind `value` (profiles/r02_launches.md)
COMPRESS_KERNEL = "snp::k_compress_v3<SNP_HASH_CRC32C, 3>"
WORKLOAD = "batched decompress: 2^20 x 64 KiB precompressed blocks per GPU (Silesia-mix synthetic), device-resident"


# ----------------------------------------------------------------------------- data

def load_corpus():
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
    get = lambda *names: np.concatenate([z["corpus/" + n] for n in names])
    return {
        "text": get("alice29.txt", "asyoulik.txt", "lcet10.txt", "plrabn12.txt"),
        "markup": get("html", "urls.10K", "geo.protodata"),
        "binary": get("kppkn.gtb"),
        "jpeg": get("fireworks.jpeg"),
    }


def make_blocks(torch, corpus_dev, first_block: int, count: int, dev, force_class: int | None = None):
    """'Silesia-mix synthetic' (SURVEY.md 8(d) config 2): deterministic per (first_block, count).
    Classes by block index: text 30 %, markup 25 %, binary 25 % (half kppkn.gtb windows, half
    LZ-synthetic: 32 fresh bytes + a 32-byte true match 1..32 KiB back), database-like records 10 %,
    incompressible 10 % (half jpeg windows, half PRNG).  Corpus windows get one byte per 4 KiB
    XOR-perturbed so that no two blocks are identical."""
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0000 + first_block)
    idx = torch.arange(first_block, first_block + count, device=dev, dtype=torch.int64)
    sel = (idx * 2654435761 >> 7) % 100
    if force_class is not None:  # diagnostics only (scratch/class_bench.py): every block from one class
        sel = torch.full_like(sel, force_class)
    out = torch.empty((count, BLOCK), dtype=torch.uint8, device=dev)

    def windows(name, rows):
        c = corpus_dev[name]
        starts = torch.randint(0, c.numel() - BLOCK, (rows.numel(),), device=dev, generator=g)
        w = c.unfold(0, BLOCK, 1).index_select(0, starts)
        r = torch.arange(rows.numel(), device=dev).unsqueeze(1)
        pos = torch.randint(0, 4096, (rows.numel(), 16), device=dev, generator=g) + \
            torch.arange(16, device=dev) * 4096
        val = torch.randint(1, 256, (rows.numel(), 16), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
        w[r, pos] = w[r, pos] ^ val
        out[rows] = w

    def pick(lo, hi):
        return torch.nonzero((sel >= lo) & (sel < hi)).squeeze(1)

    for name, lo, hi in (("text", 0, 30), ("markup", 30, 55), ("binary", 55, 67), ("jpeg", 90, 95)):
        rows = pick(lo, hi)
        if rows.numel():
            windows(name, rows)
    rows = pick(67, 80)  # LZ-synthetic: 64-byte segments = 32 fresh bytes + a 32-byte match that is a true copy of
    if rows.numel():  # the fresh run of a segment 1..32 KiB back (where the reference's probe loop has inserted
        n = rows.numel()  # every position, so its compressor finds the match: ratio ~0.59)
        R = torch.randint(0, 256, (n, BLOCK), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
        nseg = BLOCK // 64
        seg = torch.arange(nseg, device=dev)
        back = torch.randint(16, 512, (n, nseg), device=dev, generator=g)  # segments back: 1..32 KiB
        src_seg = torch.where(seg >= 16, seg - torch.minimum(back, seg.unsqueeze(0).expand(n, -1)), seg)
        k32 = torch.arange(32, device=dev)
        src = (src_seg.unsqueeze(2) * 64 + k32).view(n, nseg * 32)          # fresh run of the source segment
        v = R.view(n, nseg, 64)
        copy = R.gather(1, src).view(n, nseg, 32)
        v[:, 16:, 32:] = copy[:, 16:, :]
        out[rows] = R
    rows = pick(80, 90)  # database-like fixed-width records
    if rows.numel():
        n = rows.numel()
        nrec = BLOCK // 64
        rec = torch.full((n, nrec, 64), 0x20, dtype=torch.uint8, device=dev)
        ctr = (torch.arange(nrec, device=dev).unsqueeze(0) + idx[rows].unsqueeze(1) * 1000).to(torch.int32)
        rec[:, :, 0:4] = ctr.unsqueeze(2).bitwise_right_shift(torch.tensor([0, 8, 16, 24], device=dev, dtype=torch.int32)).to(torch.uint8)
        rec[:, :, 4:12] = torch.randint(0, 4, (n, nrec, 8), device=dev, generator=g, dtype=torch.int32).to(torch.uint8) + 0x30
        rec[:, :, 12:28] = torch.randint(0, 256, (n, 1, 16), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
        rec[:, :, 28:32] = ((ctr // 37).unsqueeze(2).bitwise_right_shift(torch.tensor([0, 8, 16, 24], device=dev, dtype=torch.int32))).to(torch.uint8)
        out[rows] = rec.view(n, BLOCK)
    rows = pick(95, 100)  # PRNG bytes
    if rows.numel():
        out[rows] = torch.randint(0, 256, (rows.numel(), BLOCK), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
    return out


def make_blocks_config3(torch, first_block: int, count: int, dev):
    """BASELINE config 3, '50 % compressible synthetic' (SURVEY.md 8(d)): 512 stripes of 128 B per
    block; each stripe = 48 fresh PRNG bytes followed by an 80-byte copy of the start of the stripe
    4096 B earlier (the first 4 KiB are all fresh).  SURVEY's 64+64 split lands at ratio 0.63-0.75
    with the reference's skip heuristic; 48+80 realises 0.51 (recorded in the JSON line)."""
    g = torch.Generator(device=dev)
    g.manual_seed(0xC0DE0000 + first_block)
    a = torch.randint(0, 256, (count, BLOCK), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
    v = a.view(count, BLOCK // 128, 128)
    for s0 in range(32, BLOCK // 128):  # sequential: the source may itself contain a copy
        v[:, s0, 48:] = v[:, s0 - 32, :80]
    return a


def block_checksums(torch, blocks_u8, weights):
    """Position-sensitive 64-bit checksum per block (wrapping int64 arithmetic)."""
    v = blocks_u8.view(torch.int64).view(-1, BLOCK // 8)
    return (v * weights).sum(dim=1)


def prepare_batch(torch, engine, n_blocks: int, first_block: int, dev, sub: int = 8192, cap_ratio: float = 0.70):
    """Generate raw blocks, compress them with the GPU compressor, keep dense compressed bytes
    + per-block offsets/lengths + raw checksums.  Raw data is not kept."""
    corpus_dev = {k: torch.from_numpy(v).to(dev) for k, v in load_corpus().items()}
    gw = torch.Generator(device=dev)
    gw.manual_seed(12345)
    weights = torch.randint(-(2**62), 2**62, (BLOCK // 8,), device=dev, generator=gw, dtype=torch.int64) | 1
    comp_cap = int(n_blocks * BLOCK * cap_ratio) + (64 << 20)
    comp = torch.empty(comp_cap, dtype=torch.uint8, device=dev)
    c_off = torch.empty(n_blocks, dtype=torch.int64, device=dev)
    c_len = torch.empty(n_blocks, dtype=torch.int32, device=dev)
    sums = torch.empty(n_blocks, dtype=torch.int64, device=dev)
    sub = min(sub, n_blocks)
    slots = torch.empty(sub * PITCH, dtype=torch.uint8, device=dev)
    s_off = torch.arange(sub, device=dev, dtype=torch.int64) * PITCH
    s_cap = torch.full((sub,), PITCH, dtype=torch.int32, device=dev)
    s_len = torch.zeros(sub, dtype=torch.int32, device=dev)
    s_st = torch.zeros(sub, dtype=torch.int32, device=dev)
    r_off = torch.arange(sub, device=dev, dtype=torch.int64) * BLOCK
    r_len = torch.full((sub,), BLOCK, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    cur = 0
    col = torch.arange(PITCH, device=dev, dtype=torch.int32)
    for b0 in range(0, n_blocks, sub):
        m = min(sub, n_blocks - b0)
        raw = make_blocks(torch, corpus_dev, first_block + b0, m, dev)
        sums[b0:b0 + m] = block_checksums(torch, raw, weights)
        engine.compress_batch_device(raw.view(-1), r_off[:m], r_len[:m], slots, s_off[:m], s_cap[:m],
                                     s_len[:m], s_st[:m], 0, stream)
        torch.cuda.synchronize()
        assert int(s_st[:m].abs().sum()) == 0, "compressor reported an error"
        lens = s_len[:m].to(torch.int64)
        total = int(lens.sum())
        assert cur + total <= comp_cap, "compressed buffer too small"
        mask = col.unsqueeze(0) < s_len[:m].unsqueeze(1)
        comp[cur:cur + total] = slots.view(sub, PITCH)[:m][mask]
        c_off[b0:b0 + m] = cur + torch.cumsum(lens, 0) - lens
        c_len[b0:b0 + m] = s_len[:m]
        cur += total
        del raw, mask
    del slots
    torch.cuda.empty_cache()
    return comp, c_off, c_len, sums, weights, cur


# --------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock + throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------ reference

def cpu_reference_run(comp_host: np.ndarray, off: np.ndarray, ln: np.ndarray, threads: int, min_seconds: float):
    """Times the oracle's multi-threaded batched decompress (the reference algorithm restated in C)."""
    from oracle import pyoracle as O
    n = len(off)
    out = np.empty(n * BLOCK, np.uint8)
    o_off = np.arange(n, dtype=np.uint64) * BLOCK
    o_cap = np.full(n, BLOCK, np.uint32)
    O.decompress_batch(comp_host, off, ln, out, o_off, o_cap, threads)  # warm
    t0 = time.perf_counter()
    passes = 0
    while True:
        bad, _, _ = O.decompress_batch(comp_host, off, ln, out, o_off, o_cap, threads)
        assert bad == 0
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return passes * n * BLOCK / dt / 1e9, passes, dt


def sanity_anchor(comp_host: np.ndarray, off: np.ndarray, ln: np.ndarray, max_blocks: int = 1024):
    """Google's C++ Snappy (pyarrow's bundled codec), one thread, same blocks: an independent yard-stick for
    the oracle port.  None when pyarrow lacks the codec."""
    try:
        import pyarrow as pa
        codec = pa.Codec("snappy")
    except Exception:
        return None
    n = min(len(off), max_blocks)
    blocks = [comp_host[int(off[i]): int(off[i]) + int(ln[i])].tobytes() for i in range(n)]
    t0 = time.perf_counter()
    for b in blocks:
        codec.decompress(b, decompressed_size=BLOCK)
    return round(n * BLOCK / (time.perf_counter() - t0) / 1e9, 3)


def dotnet_probe():
    """the real C# reference can only be timed where a .NET SDK exists."""
    import shutil
    exe = shutil.which("dotnet")
    if not exe:
        return None
    try:
        return subprocess.run([exe, "--version"], capture_output=True, text=True, timeout=20).stdout.strip() or None
    except Exception:
        return None


def host_sample_blocks(n_blocks: int, first_block: int = 0):
    """CPU-only construction of a bounded sample of the workload for --impl reference (no GPU): raw blocks of the
    config-2 mixture, compressed with the oracle.  Returns (slots, offs, lens)."""
    import torch
    from oracle import pyoracle as O
    corpus = {k: torch.from_numpy(v) for k, v in load_corpus().items()}
    raw = np.empty((n_blocks, BLOCK), np.uint8)
    for b0 in range(0, n_blocks, 8192):
        m = min(8192, n_blocks - b0)
        raw[b0:b0 + m] = make_blocks(torch, corpus, first_block + b0, m, torch.device("cpu")).numpy()
    caps = np.full(n_blocks, PITCH, np.uint32)
    offs = np.arange(n_blocks, dtype=np.uint64) * PITCH
    slots = np.empty(n_blocks * PITCH, np.uint8)
    bad, lens, _ = O.compress_batch(raw.reshape(-1), np.arange(n_blocks, dtype=np.uint64) * BLOCK,
                                    np.full(n_blocks, BLOCK, np.uint32), slots, offs, caps, 0, os.cpu_count() or 1)
    assert bad == 0
    return slots, offs, lens


def run_reference(args):
    """The reference arm: the reference's CPU implementation of the path (the oracle C port of Snappier's algorithm --
    no .NET in this image) on all host cores, on a bounded sample of the same workloads: decompress (config 2, the
    line's value) and compress (config 3, the `compress` object)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import pyoracle as O
    O.build()
    threads = os.cpu_count() or 1
    n = args.ref_blocks
    slots, offs, lens = host_sample_blocks(n)
    out = np.empty(n * BLOCK, np.uint8)
    o_off = np.arange(n, dtype=np.uint64) * BLOCK
    o_cap = np.full(n, BLOCK, np.uint32)
    for _ in range(args.warmup):
        O.decompress_batch(slots, offs, lens, out, o_off, o_cap, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bad, _, _ = O.decompress_batch(slots, offs, lens, out, o_off, o_cap, threads)
        assert bad == 0
    dt = time.perf_counter() - t0
    val = args.steps * n * BLOCK / dt / 1e9
    del slots, out
    # compress arm: config-3 blocks
    nc = min(n, args.ref_compress_blocks)
    raw = np.empty((nc, BLOCK), np.uint8)
    for b0 in range(0, nc, 8192):
        m = min(8192, nc - b0)
        raw[b0:b0 + m] = make_blocks_config3(torch, b0, m, torch.device("cpu")).numpy()
    raw = raw.reshape(-1)
    cslots = np.empty(nc * PITCH, np.uint8)
    r_off = np.arange(nc, dtype=np.uint64) * BLOCK
    r_len = np.full(nc, BLOCK, np.uint32)
    s_off = np.arange(nc, dtype=np.uint64) * PITCH
    s_cap = np.full(nc, PITCH, np.uint32)
    for _ in range(args.warmup):
        O.compress_batch(raw, r_off, r_len, cslots, s_off, s_cap, 0, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bad, clens, _ = O.compress_batch(raw, r_off, r_len, cslots, s_off, s_cap, 0, threads)
        assert bad == 0
    cdt = time.perf_counter() - t0
    cval = args.steps * nc * BLOCK / cdt / 1e9
    how = f"{threads} pthreads, oracle C port of the reference (no .NET in this image)"
    sample = f"{n} blocks of the same synthetic mix per step, {how}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_blocks": n},
        "cpu_baseline": {"value": round(val, 3), "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "compress": {"metric": "uncompressed GB/s (batched compress, 64 KiB blocks)", "value": round(cval, 3), "unit": "GB/s",
                     "ms_per_step": round(cdt / args.steps * 1e3, 3),
                     "config": {"workload": "batched compress: 64 KiB raw blocks (50 % compressible synthetic)", "sample_blocks": nc,
                                "ratio": round(float(clens.astype(np.int64).sum()) / (nc * BLOCK), 4), "hash_mode": "crc32c"},
                     "cpu_baseline": {"value": round(cval, 3), "unit": "GB/s", "cores": threads, "kind": "port",
                                      "sample": f"{nc} config-3 blocks per step, {how}"},
                     "e2e": {"value": round(cval, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "gpu_launches": 0,
    }))


def hbm_peak():
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        return float(json.load(open(pp))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_per_block(key):
    """DRAM bytes per block of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json)."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return None, None
    j = json.load(open(tp))
    return j.get(key), j.get("source")


def cpu_compress_run(raw_host: np.ndarray, n: int, threads: int, min_seconds: float):
    """Times the oracle's multi-threaded batched compress (CRC32C hash mode) on n raw blocks held on the host."""
    from oracle import pyoracle as O
    slots = np.empty(n * PITCH, np.uint8)
    r_off = np.arange(n, dtype=np.uint64) * BLOCK
    r_len = np.full(n, BLOCK, np.uint32)
    s_off = np.arange(n, dtype=np.uint64) * PITCH
    s_cap = np.full(n, PITCH, np.uint32)
    O.compress_batch(raw_host, r_off, r_len, slots, s_off, s_cap, 0, threads)  # warm
    t0 = time.perf_counter()
    passes = 0
    while True:
        bad, lens, _ = O.compress_batch(raw_host, r_off, r_len, slots, s_off, s_cap, 0, threads)
        assert bad == 0
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return passes * n * BLOCK / dt / 1e9, passes, dt


def compress_section(args, torch, dist, engine, world, rank, local, dev):
    """BASELINE config 3: batched compress of 2^20 raw 64 KiB blocks per GPU ('50 % compressible synthetic').  The raw
    blocks are resident in HBM; the worst-case output slots (76 496 B each) of 2^20 blocks would need another 80 GB, so
    the step STREAMS them: the batch goes through the C ABI in launches of 2^19 blocks that reuse one slot buffer (the
    consumer of a real pipeline would drain it in between).  Returns rank 0's dict (None elsewhere)."""
    n = args.blocks
    chunk = min(n, 1 << 19)
    nch = (n + chunk - 1) // chunk
    raw = torch.empty((n, BLOCK), dtype=torch.uint8, device=dev)
    for b0 in range(0, n, 8192):
        m = min(8192, n - b0)
        raw[b0:b0 + m] = make_blocks_config3(torch, rank * n + b0, m, dev)
    slots = torch.empty(chunk * PITCH, dtype=torch.uint8, device=dev)
    idx = torch.arange(chunk, device=dev, dtype=torch.int64)
    r_off, s_off = idx * BLOCK, idx * PITCH
    r_len = torch.full((chunk,), BLOCK, dtype=torch.int32, device=dev)
    s_cap = torch.full((chunk,), PITCH, dtype=torch.int32, device=dev)
    s_len = torch.zeros((nch, chunk), dtype=torch.int32, device=dev)
    st = torch.zeros((nch, chunk), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    flat = raw.view(-1)

    def step():
        for c in range(nch):
            m = min(chunk, n - c * chunk)
            engine.compress_batch_device(flat[c * chunk * BLOCK:], r_off[:m], r_len[:m], slots, s_off[:m], s_cap[:m],
                                         s_len[c, :m], st[c, :m], 0, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = engine.launch_count
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = engine.launch_count - launches0
    assert int(st.abs().sum()) == 0, "compress reported errors"
    cbytes = int(s_len.to(torch.int64).sum())
    # the slots hold the LAST launch's blocks: decompress them and compare with the raw blocks (bit-exact round trip),
    # and compare a sample with the oracle's compressed bytes (outside the timed region)
    c0 = (nch - 1) * chunk
    m = n - c0
    back = torch.empty(m * BLOCK, dtype=torch.uint8, device=dev)
    o_len = torch.zeros(m, dtype=torch.int32, device=dev)
    st2 = torch.zeros(m, dtype=torch.int32, device=dev)
    engine.decompress_batch_device(slots, s_off[:m], s_len[nch - 1, :m], back, r_off[:m], r_len[:m], o_len, st2, stream)
    torch.cuda.synchronize()
    assert int(st2.abs().sum()) == 0 and torch.equal(back, flat[c0 * BLOCK:]), "compressed blocks do not round-trip"
    del back
    from oracle import pyoracle as O
    sl = s_len[nch - 1].cpu().numpy()
    for i in list(range(0, m, max(1, m // 64)))[:64]:
        assert slots[i * PITCH: i * PITCH + int(sl[i])].cpu().numpy().tobytes() == O.compress(raw[c0 + i].cpu().numpy().tobytes())[1], \
            "compressed bytes differ from the oracle"

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    b = torch.tensor([float(n) * BLOCK, float(cbytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)

    # ---- e2e through the C ABI with pinned HOST buffers (H2D + kernel + D2H timed) ----
    ne = min(args.e2e_blocks, n)
    h_raw = torch.empty(ne * BLOCK, dtype=torch.uint8).pin_memory()
    h_raw.copy_(flat[:ne * BLOCK])
    h_slots = torch.empty(ne * PITCH, dtype=torch.uint8).pin_memory()
    hr_off = np.arange(ne, dtype=np.uint64) * BLOCK
    hr_len = np.full(ne, BLOCK, np.uint32)
    hs_off = np.arange(ne, dtype=np.uint64) * PITCH
    hs_cap = np.full(ne, PITCH, np.uint32)
    np_raw, np_slots = h_raw.numpy(), h_slots.numpy()
    for _ in range(2):
        engine.compress_batch_host(np_raw, hr_off, hr_len, np_slots, hs_off, hs_cap, 0)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        ol, est = engine.compress_batch_host(np_raw, hr_off, hr_len, np_slots, hs_off, hs_cap, 0)
    torch.cuda.synchronize()
    e_dt = (time.perf_counter() - t0) / e2e_steps
    assert not est.any() and np.array_equal(ol, s_len[0, :ne].cpu().numpy().astype(ol.dtype))
    te = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * ne * BLOCK / float(te[0]) / 1e9
    e_cbytes = int(ol.astype(np.int64).sum())
    if rank != 0:
        return None
    ms = float(t[0])
    u, cb = float(b[0]), float(b[1])
    peak, peak_src = hbm_peak()
    alg = float(n) * BLOCK + float(cbytes)  # rank 0's step: U read + C written
    per_block, tsrc = traffic_per_block("compress_dram_bytes_per_block")
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        nc = min(args.cpu_blocks, n)
        threads = os.cpu_count() or 1
        v, passes, dt = cpu_compress_run(flat[:nc * BLOCK].cpu().numpy(), nc, threads, args.cpu_seconds)
        cpu = {"value": round(v, 3), "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": f"first {nc} blocks of rank 0's batch, {passes} passes in {dt:.1f} s, oracle C port of the reference "
                         "compressor (CRC32C hash mode; Snappier's C# cannot run here)"}
    return {
        "metric": "uncompressed GB/s (batched compress, 64 KiB blocks)", "value": round(u / ms / 1e6, 2), "unit": "GB/s",
        "ms_per_step": round(ms, 4),
        "config": {"workload": f"batched compress: {n} x 64 KiB raw blocks per GPU (50 % compressible synthetic), device-resident, "
                               f"streamed as {nch} launches of {chunk} blocks through one slot buffer",
                   "ratio": round(cb / u, 4), "hash_mode": "crc32c"},
        "roofline": {"bound": "hbm", "achieved": round(alg / ms / 1e6, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(alg / ms / 1e6 / peak, 4), "traffic": int(per_block * n) if per_block else None,
                     "traffic_source": tsrc, "peak_source": peak_src, "kernel": COMPRESS_KERNEL,
                     "algorithmic_bytes_per_step": int(alg)},
        "cpu_baseline": cpu,
        "e2e": {"value": round(e2e_val, 3), "unit": "GB/s", "h2d_bytes_per_step": int(ne * BLOCK + ne * 24),
                "d2h_bytes_per_step": int(e_cbytes + ne * 8), "blocks_per_step": ne,
                "path": "snp_compress_batch(SNP_MEM_HOST) on pinned host buffers"},
        "gpu_launches": int(launches), "clocks": clk.summary()}


def run_compress(args, torch, dist, engine, world, rank, local, dev):
    """--workload compress: config 3 as a line of its own."""
    sec = compress_section(args, torch, dist, engine, world, rank, local, dev)
    if rank == 0:
        line = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic"}
        line.update(sec)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_frame(args, torch, engine, rank, local, dev):
    """BASELINE config 4: the SnappyStream framing format end to end (stream id + 64 KiB chunks with masked
    CRC32C, raw-vs-compressed choice) over a synthetic stream of the config-2 mixture.  Both calls take
    HOST buffers, so these are e2e numbers (H2D + kernels + D2H inside the timed region).  Extra line."""
    import ctypes as C
    from snappier_b200 import _native as N
    if rank != 0:
        return
    n_blocks = int(args.frame_gib * (1 << 30)) // BLOCK
    corpus_dev = {k: torch.from_numpy(v).to(dev) for k, v in load_corpus().items()}
    raw = torch.empty(n_blocks * BLOCK, dtype=torch.uint8).pin_memory()
    for b0 in range(0, n_blocks, 8192):
        m = min(8192, n_blocks - b0)
        raw[b0 * BLOCK:(b0 + m) * BLOCK].copy_(make_blocks(torch, corpus_dev, b0, m, dev).view(-1))
    torch.cuda.synchronize()
    L = N.lib()
    cap = int(L.snp_frame_max_compressed_length(raw.numel()))
    framed = torch.empty(cap, dtype=torch.uint8).pin_memory()
    back = torch.empty(raw.numel(), dtype=torch.uint8).pin_memory()
    w = C.c_size_t(0)

    def compress():
        st = L.snp_frame_compress(raw.data_ptr(), raw.numel(), framed.data_ptr(), cap, C.byref(w), 0)
        assert st == 0, st
        return w.value

    def decompress(nbytes):
        st = L.snp_frame_decompress(framed.data_ptr(), nbytes, back.data_ptr(), back.numel(), C.byref(w))
        assert st == 0 and w.value == raw.numel(), (st, w.value)

    launches0 = engine.launch_count
    nbytes = compress()  # warm-up (allocates the staging buffers)
    decompress(nbytes)
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            nbytes = compress()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            decompress(nbytes)
        t2 = time.perf_counter()
    assert torch.equal(back, raw), "framed round trip differs"
    # parity of the first chunks against the oracle's framing (outside the timed region)
    from oracle import pyoracle as O
    k = 8 * BLOCK
    want = O.frame_compress(raw[:k].numpy().tobytes())
    assert framed[:len(want)].numpy().tobytes() == want, "framed bytes differ from the oracle"
    cg = raw.numel() * args.steps / (t1 - t0) / 1e9
    dg = raw.numel() * args.steps / (t2 - t1) / 1e9
    print(json.dumps({
        "metric": "uncompressed GB/s (SnappyStream framing format, host buffers end to end)",
        "value": round(2 / (1 / cg + 1 / dg), 3), "unit": "GB/s", "n_gpus": 1, "steps": args.steps, "warmup": 1,
        "ms_per_step": round((t2 - t0) / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"framed stream: {args.frame_gib} GiB of the config-2 mixture in 64 KiB chunks, one whole-stream "
                               "snp_frame_compress + snp_frame_decompress per step (pinned host buffers)",
                   "framed_bytes": int(nbytes), "ratio": round(nbytes / raw.numel(), 4), "hash_mode": "crc32c"},
        "frame_compress_GBps": round(cg, 3), "frame_decompress_GBps": round(dg, 3),
        "e2e": {"value": round(2 / (1 / cg + 1 / dg), 3), "unit": "GB/s", "h2d_bytes_per_step": int(raw.numel() + nbytes),
                "d2h_bytes_per_step": int(raw.numel() + nbytes)},
        "gpu_launches": "n/a (the library's default context)" if engine.launch_count == launches0 else int(engine.launch_count - launches0),
        "clocks": clk.summary()}))


def roundtrip_section(args, torch, dist, engine, world, rank, local, dev):
    """BASELINE config 5: mixed-block corpus (config-2 mixture and config-3 blocks 1:1), 2^17 blocks = 8 GiB per GPU
    (64 GiB at 8 GPUs), sharded by contiguous block range, compress then decompress, verified by per-block checksums.
    Two measurements:
      * born sharded: every rank holds its shard already; no data-path collective (`value`);
      * root-sourced (N > 1): rank 0 holds the whole corpus; ONE scatter of the raw block ranges over NCCL, compress,
        decompress, ONE gather(v) of the packed compressed blocks back to rank 0 -- bound by rank 0's NVLink egress /
        ingress, reported per phase (`nccl_scatter_gather`).
    Returns rank 0's dict (None elsewhere)."""
    n = min(args.blocks, 1 << 17)
    corpus_dev = {k: torch.from_numpy(v).to(dev) for k, v in load_corpus().items()}

    def gen(first, count):
        out = torch.empty((count, BLOCK), dtype=torch.uint8, device=dev)
        for b0 in range(0, count, 8192):
            m = min(8192, count - b0)
            a = make_blocks(torch, corpus_dev, first + b0, m, dev)
            c3 = make_blocks_config3(torch, first + b0, m, dev)
            odd = (torch.arange(m, device=dev) % 2 == 1)
            a[odd] = c3[odd]
            out[b0:b0 + m] = a
        return out

    raw = gen(rank * n, n)
    gw = torch.Generator(device=dev)
    gw.manual_seed(12345)
    weights = torch.randint(-(2**62), 2**62, (BLOCK // 8,), device=dev, generator=gw, dtype=torch.int64) | 1
    sums = block_checksums(torch, raw.view(-1), weights)
    slots = torch.empty(n * PITCH, dtype=torch.uint8, device=dev)
    back = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    r_off, s_off = idx * BLOCK, idx * PITCH
    r_len = torch.full((n,), BLOCK, dtype=torch.int32, device=dev)
    s_cap = torch.full((n,), PITCH, dtype=torch.int32, device=dev)
    s_len = torch.zeros(n, dtype=torch.int32, device=dev)
    o_len = torch.zeros(n, dtype=torch.int32, device=dev)
    st1 = torch.zeros(n, dtype=torch.int32, device=dev)
    st2 = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step(src=None):
        engine.compress_batch_device(raw.view(-1) if src is None else src, r_off, r_len, slots, s_off, s_cap, s_len, st1, 0, stream)
        engine.decompress_batch_device(slots, s_off, s_len, back, r_off, r_len, o_len, st2, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = engine.launch_count
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = engine.launch_count - launches0
    assert int(st1.abs().sum()) == 0 and int(st2.abs().sum()) == 0
    assert torch.equal(block_checksums(torch, back, weights), sums), "round trip differs from the source"
    cbytes = float(s_len.to(torch.int64).sum())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    b = torch.tensor([float(n) * BLOCK, cbytes], dtype=torch.float64, device=dev)
    link = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        # ---- root-sourced: the whole corpus (world x n blocks) starts and ends on rank 0 ----
        from snappier_b200 import sharding as SH
        del raw
        torch.cuda.empty_cache()
        tot = n * world
        if rank == 0:
            src = gen(0, tot).view(-1)
            all_sums = block_checksums(torch, src, weights)
            soff = torch.arange(tot, device=dev, dtype=torch.int64) * BLOCK
            slen = torch.full((tot,), BLOCK, dtype=torch.int32, device=dev)
        else:
            src = soff = slen = None
        ph = {}
        g_base = g_off = g_len = None
        for it in range(2):  # the first pass pays NCCL's lazy peer-to-peer connection setup: the second is reported
            g_base = g_off = g_len = None  # the caching allocator hands the first pass's buffers to the second
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            barrier()
            ev[0].record()
            my_base, my_off, my_len, first, _ = SH.scatter_batch(src, soff, slen, 0, dev)
            ev[1].record()
            engine.compress_batch_device(my_base, my_off, my_len, slots, s_off, s_cap, s_len, st1, 0, stream)
            ev[2].record()
            engine.decompress_batch_device(slots, s_off, s_len, back, r_off, r_len, o_len, st2, stream)
            ev[3].record()
            g_base, g_off, g_len = SH.gather_batch(slots, s_off, s_len, 0, engine=engine)
            ev[4].record()
            barrier()
            ph = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(4)], dtype=torch.float64, device=dev)
            dist.all_reduce(ph, op=dist.ReduceOp.MAX)
            # every rank's decompressed shard equals the slice of the corpus it was sent
            assert int(st1.abs().sum()) == 0 and int(st2.abs().sum()) == 0
            assert torch.equal(back[: my_base.numel()], my_base[: back.numel()]), "root-sourced round trip differs"
            del my_base
        sc_ms, c_ms, d_ms, g_ms = [float(x) for x in ph]
        if rank == 0:
            # the gathered batch decodes to the corpus: checksum a sample of it (first blocks of every rank's range)
            gl = g_len.to(torch.int64)
            g_bytes = int(gl.sum())
            ns = min(n, 4096)
            for r in range(world):
                sel = slice(r * n, r * n + ns)
                tmp = torch.empty(ns * BLOCK, dtype=torch.uint8, device=dev)
                engine.decompress_batch_device(g_base, g_off[sel].contiguous(), g_len[sel].contiguous(), tmp, r_off[:ns], r_len[:ns],
                                               o_len[:ns], st2[:ns], stream)
                torch.cuda.synchronize()
                assert int(st2[:ns].abs().sum()) == 0 and torch.equal(block_checksums(torch, tmp, weights), all_sums[sel])
            total_ms = sc_ms + c_ms + d_ms + g_ms
            link = {"blocks_per_rank": n, "corpus_bytes": int(tot) * BLOCK,
                    "scatter_raw_GBps": round(tot * BLOCK / sc_ms / 1e6, 1),
                    "gather_compressed_GBps": round(g_bytes / g_ms / 1e6, 1),
                    "scatter_ms": round(sc_ms, 2), "compress_ms": round(c_ms, 2), "decompress_ms": round(d_ms, 2),
                    "gather_ms": round(g_ms, 2),
                    "root_sourced_roundtrip_GBps": round(tot * BLOCK / total_ms / 1e6, 1),
                    "bound": f"rank 0's NVLink: it sends {world - 1}/{world} of the raw corpus and receives {world - 1}/{world} "
                             "of the compressed corpus (770 GB/s per direction measured, 900 nominal)",
                    "collectives": "1 broadcast (byte counts) + 1 grouped send/recv (scatter); 1 all_gather (counts) + "
                                   "1 grouped send/recv (gather); snp_pack_batch packs the slots before the send"}
    if rank != 0:
        return None
    ms = float(t[0])
    u, c = float(b[0]), float(b[1])
    return {
        "metric": "uncompressed GB/s (compress + decompress round trip, 64 KiB blocks)", "value": round(u / ms / 1e6, 2),
        "unit": "GB/s", "ms_per_step": round(ms, 3),
        "config": {"workload": f"round trip: {n} mixed 64 KiB blocks per GPU (config-2 mixture and config-3 blocks 1:1), "
                               "born sharded by contiguous block range, compress then decompress, checksum-verified",
                   "ratio": round(c / u, 4), "hash_mode": "crc32c", "parallelism": f"block-range shard x{world}"},
        "nccl_scatter_gather": link, "gpu_launches": int(launches), "clocks": clk.summary()}


def run_roundtrip(args, torch, dist, engine, world, rank, local, dev):
    """--workload roundtrip: config 5 as a line of its own."""
    sec = roundtrip_section(args, torch, dist, engine, world, rank, local, dev)
    if rank == 0:
        line = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic"}
        line.update(sec)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=1 << 20, help="blocks per GPU (default 2^20 = BASELINE config 2)")
    ap.add_argument("--e2e-blocks", type=int, default=1 << 16)
    ap.add_argument("--cpu-blocks", type=int, default=1 << 16)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--ref-blocks", type=int, default=1 << 16)
    ap.add_argument("--ref-compress-blocks", type=int, default=1 << 15)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decompress-only", action="store_true", help="skip the compress (and, for N > 1, round-trip) sections")
    ap.add_argument("--workload", default="decompress", choices=["decompress", "compress", "frame", "roundtrip"],
                    help="decompress = BASELINE config 2 (the headline); compress = config 3, frame = config 4, "
                         "roundtrip = config 5 (extra lines, not the driver's)")
    ap.add_argument("--frame-gib", type=float, default=16.0, help="frame workload: stream size (config 4 names 16 GiB; needs 3x that of pinned host memory)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from snappier_b200.batch import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # N > 1: every rank works from the CPUs / memory of the NUMA node its GPU hangs off (the e2e legs copy from and to
    # pinned host buffers; unbound, the ranks' PCIe streams cross the socket interconnect).  Not at N = 1: the host-core
    # baseline of that run must see all the cores.
    numa_info = None
    if world > 1 and os.environ.get("SNP_BENCH_NUMA", "1") != "0":
        from snappier_b200 import numa
        numa_info = numa.bind_to_gpu_node(local)
    engine = Engine(local)
    n = args.blocks
    if args.workload == "compress":
        return run_compress(args, torch, dist, engine, world, rank, local, dev)
    if args.workload == "frame":
        return run_frame(args, torch, engine, rank, local, dev)
    if args.workload == "roundtrip":
        return run_roundtrip(args, torch, dist, engine, world, rank, local, dev)

    comp, c_off, c_len, sums, weights, comp_bytes = prepare_batch(torch, engine, n, rank * n, dev)
    out = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
    o_off = torch.arange(n, device=dev, dtype=torch.int64) * BLOCK
    o_cap = torch.full((n,), BLOCK, dtype=torch.int32, device=dev)
    o_len = torch.zeros(n, dtype=torch.int32, device=dev)
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        engine.decompress_batch_device(comp, c_off, c_len, out, o_off, o_cap, o_len, status, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = engine.launch_count
    with ClockSampler(local) as clk:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            step()
            ev[i + 1].record()
        barrier()
    launches = engine.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]

    # verify the timed output: status, lengths, checksum of checksums
    assert int(status.abs().sum()) == 0 and bool((o_len == BLOCK).all()), "decompress reported errors"
    got = block_checksums(torch, out, weights)
    assert torch.equal(got, sums), "decompressed bytes differ from the raw blocks"

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    b = torch.tensor([float(n) * BLOCK, float(comp_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    total_ms = float(t[0])
    u_bytes, c_bytes = float(b[0]), float(b[1])
    value = u_bytes * args.steps / (total_ms * 1e-3) / 1e9

    # ---- e2e through the C ABI with pinned HOST buffers (H2D + kernel + D2H timed) ----
    ne = min(args.e2e_blocks, n)
    e_bytes = int(c_off[ne - 1] + c_len[ne - 1]) if ne else 0
    h_in = torch.empty(e_bytes, dtype=torch.uint8).pin_memory()
    h_in.copy_(comp[:e_bytes])
    h_out = torch.empty(ne * BLOCK, dtype=torch.uint8).pin_memory()
    h_coff = c_off[:ne].cpu().numpy().astype(np.uint64)
    h_clen = c_len[:ne].cpu().numpy().astype(np.uint32)
    h_ooff = np.arange(ne, dtype=np.uint64) * BLOCK
    h_ocap = np.full(ne, BLOCK, np.uint32)
    np_in, np_out = h_in.numpy(), h_out.numpy()
    for _ in range(2):
        engine.decompress_batch_host(np_in, h_coff, h_clen, np_out, h_ooff, h_ocap)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        ol, st = engine.decompress_batch_host(np_in, h_coff, h_clen, np_out, h_ooff, h_ocap)
    torch.cuda.synchronize()
    e_dt = (time.perf_counter() - t0) / e2e_steps
    assert not st.any()
    assert torch.equal(block_checksums(torch, h_out.to(dev), weights), sums[:ne])
    te = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * ne * BLOCK / float(te[0]) / 1e9
    meta_bytes = ne * (8 + 8 + 4 + 4)

    # ---- the GPU-consumer case: compressed bytes come from the host, the decompressed blocks STAY on the device --------
    # (H2D of C only: about half the PCIe bytes of the host-to-host call.)  Two streams, chunks of 4096 blocks: the copy
    # of chunk k+1 overlaps the kernel of chunk k; device-mode batch calls of the same C ABI.
    d_in2 = torch.empty(e_bytes, dtype=torch.uint8, device=dev)
    d_out2 = torch.empty(ne * BLOCK, dtype=torch.uint8, device=dev)
    cs, ks = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ck = 4096
    bounds = [(a, min(a + ck, ne)) for a in range(0, ne, ck)]
    h_coff_i = h_coff.astype(np.int64)
    spans = [(int(h_coff_i[a]), int(h_coff_i[b - 1]) + int(h_clen[b - 1])) for a, b in bounds]

    def device_out_step():
        evs = []
        for (a, b), (lo, hi) in zip(bounds, spans):
            with torch.cuda.stream(cs):
                d_in2[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            ks.wait_event(ev)
            with torch.cuda.stream(ks):
                engine.decompress_batch_device(d_in2, c_off[a:b], c_len[a:b], d_out2, o_off[a:b], o_cap[a:b], o_len[a:b],
                                               status[a:b], ks.cuda_stream)
            evs.append(ev)
        ks.synchronize()

    for _ in range(2):
        device_out_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        device_out_step()
    torch.cuda.synchronize()
    d_dt = (time.perf_counter() - t0) / e2e_steps
    assert int(status[:ne].abs().sum()) == 0
    assert torch.equal(block_checksums(torch, d_out2, weights), sums[:ne])
    td = torch.tensor([d_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
    e2e_devout = world * ne * BLOCK / float(td[0]) / 1e9
    del d_in2, d_out2

    # ---- roofline of the (single) dominant kernel --------------------------------------
    line = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        k_ms = float(np.mean(kernel_ms))
        alg_bytes = (float(n) * BLOCK + float(comp_bytes))  # rank 0's launch: C_i read + U_i written
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        per_block, tsrc = traffic_per_block("decompress_dram_bytes_per_block")
        traffic = int(per_block * n) if per_block else None  # ncu --set full capture, scaled to this launch

        cpu = None
        if not args.no_cpu_baseline and world == 1:  # the host-core baseline is reported by the single-GPU run only
            nc = min(args.cpu_blocks, n)
            cb = int(c_off[nc - 1] + c_len[nc - 1])
            threads = os.cpu_count() or 1
            h_comp = comp[:cb].cpu().numpy()
            h_off = c_off[:nc].cpu().numpy().astype(np.uint64)
            h_len = c_len[:nc].cpu().numpy().astype(np.uint32)
            v, passes, dt = cpu_reference_run(h_comp, h_off, h_len, threads, args.cpu_seconds)
            dn = dotnet_probe()
            cpu = {"value": round(v, 3), "unit": "GB/s", "cores": threads, "kind": "port",
                   "sample": f"first {nc} blocks of rank 0's batch, {passes} passes in {dt:.1f} s, oracle C port of the "
                             f"reference algorithm (Snappier's C# cannot run here: " +
                             (f"dotnet {dn} found, see csharp/Bench)" if dn else "no .NET SDK)"),
                   "google_snappy_1thread_GBps": sanity_anchor(h_comp, h_off, h_len), "dotnet": dn}
            del h_comp

        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD if n == (1 << 20) else WORKLOAD.replace("2^20", str(n)),
                       "blocks_per_gpu": n, "block_bytes": BLOCK, "compressed_bytes_per_gpu": int(comp_bytes),
                       "ratio": round(c_bytes / u_bytes, 4), "hash_mode": "crc32c", "parallelism": f"block-range shard x{world}",
                       "l2": "inputs+outputs per step far exceed the 126 MB L2 (no flush needed)" if n * BLOCK > (1 << 30)
                             else "WARNING: working set is small relative to L2"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
                         "kernel": DECOMPRESS_KERNEL, "kernel_ms": round(k_ms, 4),
                         "algorithmic_bytes_per_launch": int(alg_bytes)},
            "cpu_baseline": cpu,
            "e2e": {"value": round(e2e_val, 3), "unit": "GB/s", "h2d_bytes_per_step": int(e_bytes + meta_bytes),
                    "d2h_bytes_per_step": int(ne * BLOCK + ne * 8), "blocks_per_step": ne,
                    "path": "snp_decompress_batch(SNP_MEM_HOST) on pinned host buffers",
                    "compressed_in_device_out": {"value": round(e2e_devout, 3), "unit": "GB/s", "h2d_bytes_per_step": int(e_bytes),
                                                 "d2h_bytes_per_step": 0,
                                                 "path": "pinned compressed bytes -> cudaMemcpyAsync -> snp_decompress_batch(SNP_MEM_DEVICE), "
                                                         "4096-block chunks on two streams; the output stays in HBM"}},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        if numa_info is not None:
            line["numa"] = numa_info

    # ---- the other half of BASELINE's metric (config 3) and, across GPUs, config 5: same JSON line ----
    del comp, out, h_in, h_out, np_in, np_out
    torch.cuda.empty_cache()
    if not args.decompress_only:
        sec = compress_section(args, torch, dist, engine, world, rank, local, dev)
        if rank == 0:
            line["compress"] = sec
            line["gpu_launches"] += sec["gpu_launches"]
        torch.cuda.empty_cache()
        if world > 1:
            rt = roundtrip_section(args, torch, dist, engine, world, rank, local, dev)
            if rank == 0:
                line["roundtrip"] = rt
                line["nccl_scatter_gather"] = rt["nccl_scatter_gather"]
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
