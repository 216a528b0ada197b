#!/usr/bin/env python3
"""Annotates `cuobjdump -sass` output of one kernel with the scheduling control fields of each instruction
(sm_70+ 128-bit encoding: stall count, yield, write/read scoreboard index, wait mask) -- shows where ptxas waits for
outstanding loads.   python tools/sass_ctrl.py <lib.so> <kernel-name-substring> [grep-regex]"""
import re
import subprocess
import sys

so, name = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
on = False
prev = None
for line in txt:
    if "Function :" in line:
        on = name in line
        continue
    if not on:
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", line)
    if m:
        prev = (m.group(1), m.group(2).strip())
        continue
    m2 = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", line)
    if m2 and prev:
        hi = int(m2.group(1), 16)
        ctrl = (hi >> 41) & 0x1fffff
        stall, yld, wbar, rbar, wait = ctrl & 15, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 63
        s = "%s  %-78s st=%2d %s w=%s r=%s wait=%s" % (prev[0], prev[1][:78], stall, "Y" if not yld else " ",
                                                    "-" if wbar == 7 else wbar, "-" if rbar == 7 else rbar,
                                                    "".join(str(i) for i in range(6) if wait >> i & 1) or "-")
        if pat is None or pat.search(s):
            print(s)
        prev = None
