#!/usr/bin/env python3
"""The B200's random-access read ceiling: 16-byte loads at pseudo-random addresses of a buffer that fits L2
(64 MiB) or does not (8 GiB), at several occupancies.  Prints G accesses/s and the DRAM-side GB/s they imply.

    python tools/random_access_probe.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from snappier_b200 import _native as N  # noqa: E402
from snappier_b200.batch import Engine  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
eng = Engine(0)
L = N.lib()
sink = torch.zeros(4, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
sms = torch.cuda.get_device_properties(0).multi_processor_count
res = []
for span_mib in (64, 1024, 8192):
    buf = torch.empty(span_mib << 20, dtype=torch.uint8, device=dev)
    buf.view(torch.int64).random_()
    for ctas in (1, 2, 4, 8):
        reads = 4096
        for _ in range(2):
            L.snp_diag_random_reads(eng._ctx, buf.data_ptr(), buf.numel(), ctas, reads, sink.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.snp_diag_random_reads(eng._ctx, buf.data_ptr(), buf.numel(), ctas, reads, sink.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        n = sms * ctas * 256 * reads
        row = {"span_MiB": span_mib, "threads_per_SM": ctas * 256, "G_accesses_per_s": round(n / ms / 1e6, 2),
               "GBps_at_32B_sectors": round(n * 32 / ms / 1e6, 1), "GBps_at_64B": round(n * 64 / ms / 1e6, 1),
               "GBps_at_128B_lines": round(n * 128 / ms / 1e6, 1)}
        res.append(row)
        print(json.dumps(row), flush=True)
    del buf
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r02_random_access_probe.json"), "w"), indent=1)
