"""Host-side check of the lane-per-block decompress engine (k_decompress_v8, the large-batch default): its lane function
compiled with g++ against tests/cpp/simt_emu.h, ONE emulated warp taking the whole batch through the work counter (every
lane its own block, at its own misalignment), compared with the oracle (status, length, bytes, guard bytes around every
output).  cp.async is synchronous in the emulator, so this checks the ring / accumulator / flush / scheduling LOGIC; the
GPU parity tests remain the proof for the compiled kernel."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H
from tests.helpers import BUILD, ROOT, handmade_tag_forms


@pytest.fixture(scope="module")
def emu8():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "emu_v8")
    srcs = [os.path.join(ROOT, "tests", "cpp", "emu_v8.cpp"), os.path.join(ROOT, "tests", "cpp", "simt_emu.h")] + [
        os.path.join(ROOT, "snappier_b200", "csrc", f) for f in
        ("snp_decompress_v8.cuh", "snp_decompress_v7.cuh", "snp_tma.cuh", "snp_decompress_v1.cuh", "snp_common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, srcs[0]])
    return exe


@pytest.mark.parametrize("oring", [128, 256, 1282, 1284, 1281])
def test_emu_v8_blocks(oracle, fixtures, kats, emu8, tmp_path, oring):
    items = [oracle.compress(s)[1] for s in H.edge_strings(kats)]
    items += [oracle.compress(b)[1] for b in (b"", b"a", b"abc" * 100, b"\x00" * 65536, b"ab" * 700 + b"c" * 3000)]
    items += H.bad_blocks(oracle, fixtures)
    for name in ("alice29.txt", "html", "kppkn.gtb", "fireworks.jpeg", "geo.protodata", "urls.10K"):
        blocks = H.blocks_of(fixtures[f"corpus/{name}"])
        items += [oracle.compress(blocks[0])[1], oracle.compress(blocks[-1])[1]]
    items += [oracle.compress(b)[1] for b in H.synthetic_blocks(5, 6)]
    H.emu_check(oracle, emu8, items, tmp_path, oring, seed=oring)


def test_emu_v8_handmade_and_fuzz(oracle, emu8, tmp_path):
    items = handmade_tag_forms()
    rng = np.random.default_rng(8)
    base_blocks = [oracle.compress(b)[1] for b in H.synthetic_blocks(77, 12, size=4096)]
    for i in range(300):
        b = bytearray(base_blocks[i % len(base_blocks)])
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        if i % 5 == 0:
            b = b[: int(rng.integers(0, len(b)))]
        items.append(bytes(b))
    H.emu_check(oracle, emu8, items, tmp_path, 128, seed=9)


def test_emu_v8_big_blocks_ragged_and_long_literals(oracle, fixtures, emu8, tmp_path):
    """Blocks above 1 MiB (handed to the whole warp), blocks above 64 KiB under one header, ragged sizes around the ring
    and chunk sizes, literal runs around the bulk-copy threshold at every alignment, short-period copies."""
    rng = np.random.default_rng(3)
    data = fixtures["corpus/html_x_4"][:300000]
    items = [oracle.compress(data)[1], oracle.compress(fixtures["corpus/alice29.txt"])[1],
             oracle.compress((fixtures["corpus/html_x_4"] * 3)[:1200000])[1]]
    for n in (1, 7, 8, 9, 14, 15, 16, 17, 47, 48, 49, 60, 61, 62, 63, 64, 65, 127, 128, 129, 239, 240, 241, 255, 256, 257,
              271, 272, 273, 511, 512, 513, 1023, 1024, 1025, 4095, 4096, 4097):
        items.append(oracle.compress(rng.integers(0, 256, size=n, dtype=np.uint8).tobytes())[1])
        items.append(oracle.compress((bytes(rng.integers(97, 101, size=7, dtype=np.uint8)) * (n // 7 + 1))[:n])[1])
    for n in range(230, 300, 3):  # literal runs around BULK_MIN between compressible stretches
        items.append(oracle.compress(b"x" * 100 + rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() + b"y" * 3000)[1])
    for period in range(1, 20):   # overlapping copies of every short period, then a far reference back into them
        pat = bytes(rng.integers(0, 256, size=period, dtype=np.uint8))
        items.append(oracle.compress(pat * 300 + rng.integers(0, 256, size=500, dtype=np.uint8).tobytes() + pat * 40)[1])
    H.emu_check(oracle, emu8, items, tmp_path, 128, seed=11)
