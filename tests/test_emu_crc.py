"""Host-side check of the warp-parallel CRC32C (snp_frame.cuh: 128-byte strides per lane, GF(2) alignment of the lane
states, XOR reduction) on tests/cpp/simt_emu.h against the oracle and the reference's known-answer tests."""
from __future__ import annotations

import os
import struct
import subprocess

import numpy as np

from tests.helpers import BUILD, ROOT


def test_emu_crc32c_warp(oracle, tmp_path):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "emu_crc")
    srcs = [os.path.join(ROOT, "tests", "cpp", "emu_crc.cpp"), os.path.join(ROOT, "tests", "cpp", "simt_emu.h"),
            os.path.join(ROOT, "snappier_b200", "csrc", "snp_frame.cuh"), os.path.join(ROOT, "snappier_b200", "csrc", "snp_common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, srcs[0]])
    rng = np.random.default_rng(3)
    items = [b"", b"a", b"123456789"]
    for n in (3, 4, 5, 127, 128, 129, 131, 255, 256, 257, 1000, 4096, 65535, 65536):
        items.append(rng.integers(0, 256, size=n, dtype=np.uint8).tobytes())
    blob = bytearray(struct.pack("<I", len(items)))
    for i, b in enumerate(items):
        blob += struct.pack("<II", len(b), i) + b  # skew i & 7: every alignment
    fin, fout = os.path.join(tmp_path, "crc_in.bin"), os.path.join(tmp_path, "crc_out.bin")
    with open(fin, "wb") as f:
        f.write(blob)
    subprocess.check_call([exe, fin, fout], timeout=600)
    raw = open(fout, "rb").read()
    for i, b in enumerate(items):
        crc, masked = struct.unpack_from("<II", raw, 8 * i)
        assert crc == oracle.crc32c(b), (i, len(b))
        assert masked == oracle.crc32c_masked(b), (i, len(b))
    assert struct.unpack_from("<I", raw, 8 * 2)[0] == 0xE3069283  # "123456789" (Crc32CAlgorithmTests.cs:8-11)
