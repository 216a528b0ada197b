#!/usr/bin/env python3
"""Turn ncu captures brought back in gpurun_out/ into the small, tracked summaries under profiles/.

    python profiles/summarize.py rep  gpurun_out/prof_decomp_r01.ncu-rep profiles/r01_decompress_v3_ncu.md "title"
    python profiles/summarize.py list gpurun_out/launches_r01.csv        profiles/r01_launches.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def rep(path, dest, title):
    rows = ncu_csv(path, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    L = [f"# {title}", "", f"Source capture: `{path}` (`ncu --set full --clock-control none --import-source on`, one launch).",
         "Numbers under a profiler are for shares and counters, never bench values.", "",
         f"Kernel: `{d.get('Kernel Name', ('?',))[0][:100]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in d:
            L.append(f"| `{k}` | {d[k][0]} | {d[k][1]} |")
    L += ["", "Warp-stall reasons (warps stalled per issue-active cycle):", "", "| reason | ratio |", "|---|---|"]
    st = []
    for h in hdr:
        m = re.match(STALLS.replace("%s", "(.*)"), h)
        if m:
            st.append((float(d[h][0] or 0), m.group(1)))
    for v, n in sorted(st, reverse=True)[:8]:
        L.append(f"| {n} | {v:.3f} |")
    # hottest SASS by stall samples / executed count
    src = ncu_csv(path, "source", ["--print-source", "sass"])
    if len(src) > 3:
        h = src[1]
        iS, iI, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        body = [r for r in src[2:] if len(r) > iW and r[iI].isdigit()]
        tot_i = sum(int(r[iI]) for r in body)
        tot_w = sum(int(r[iW]) for r in body) or 1
        L += ["", f"Total warp-instructions executed: {tot_i:,}.  Hottest SASS lines by stall samples:", "",
              "| samples % | executed | SASS |", "|---|---|---|"]
        for r in sorted(body, key=lambda r: -int(r[iW]))[:12]:
            L.append(f"| {100 * int(r[iW]) / tot_w:.1f} | {int(r[iI]):,} | `{r[iS].strip()[:70]}` |")
    open(dest, "w").write("\n".join(L) + "\n")
    print("wrote", dest)


def launches(path, dest):
    lines = [l for l in open(path, newline="") if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    iN, iV, iU, iM = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Unit", "Metric Name"))
    scale = {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}
    agg, dec = collections.OrderedDict(), []
    for r in rd:
        if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
            continue
        ns = float(r[iV].replace(",", "")) * scale.get(r[iU], 1)
        name = re.sub(r"\(.*", "", r[iN])[:80]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        if "k_decompress" in r[iN]:
            dec.append(ns / 1e6)
    tot = sum(a[1] for a in agg.values())
    L = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (round 1)", "",
         f"Source: `{path}` (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`), "
         f"{sum(a[0] for a in agg.values())} launches, cold-cache and serialised: compare SHARES, not absolutes.", "",
         "The timed region of bench.py contains exactly one kernel per step (`snp::k_decompress_v3`, preceded by an",
         "8-byte `cudaMemsetAsync` of its work counter); everything else below is input preparation (torch generators +",
         "our GPU compressor producing the compressed blocks) and the e2e leg, all outside the timed region.", "",
         f"`k_decompress_v3` launches (ms): {[round(x, 2) for x in dec]} -- the ~276 ms ones are the 2^20-block steps",
         "(3 warm-up + 2 timed; share of the timed region: 100 %), the ~10 ms ones are the 32768-block e2e steps.", "",
         "| launches | total ms | share of all GPU time | kernel |", "|---|---|---|---|"]
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        L.append(f"| {c} | {ns / 1e6:.2f} | {100 * ns / tot:.2f} % | `{k}` |")
    open(dest, "w").write("\n".join(L) + "\n")
    print("wrote", dest)


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        rep(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else sys.argv[2])
    else:
        launches(sys.argv[2], sys.argv[3])
