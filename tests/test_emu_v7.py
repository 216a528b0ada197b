"""Host-side check of the tag-group decompress engine (k_decompress_v7, the default): its block function compiled with
g++ against tests/cpp/simt_emu.h and compared with the oracle (status, length, bytes, guard bytes around the output) for
every window size the kernel is instantiated with.  The TMA copies are synchronous in the emulator, so this checks the
ring / advance-table / group / round / window-flush LOGIC; the GPU parity tests remain the proof for the compiled kernel."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H
from tests.helpers import BUILD, ROOT, handmade_tag_forms


@pytest.fixture(scope="module")
def emu7():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "emu_v7")
    srcs = [os.path.join(ROOT, "tests", "cpp", "emu_v7.cpp"), os.path.join(ROOT, "tests", "cpp", "simt_emu.h")] + [
        os.path.join(ROOT, "snappier_b200", "csrc", f) for f in
        ("snp_decompress_v7.cuh", "snp_tma.cuh", "snp_decompress_v1.cuh", "snp_common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, srcs[0]])
    return exe


@pytest.mark.parametrize("window", [2048, 4096, 1024])
def test_emu_v7_blocks(oracle, fixtures, kats, emu7, tmp_path, window):
    items = [oracle.compress(s)[1] for s in H.edge_strings(kats)]
    items += [oracle.compress(b)[1] for b in (b"", b"a", b"abc" * 100, b"\x00" * 65536, b"ab" * 700 + b"c" * 3000)]
    items += H.bad_blocks(oracle, fixtures)
    for name in ("alice29.txt", "html", "kppkn.gtb", "fireworks.jpeg", "geo.protodata", "urls.10K"):
        blocks = H.blocks_of(fixtures[f"corpus/{name}"])
        items += [oracle.compress(blocks[0])[1], oracle.compress(blocks[-1])[1]]
    items += [oracle.compress(b)[1] for b in H.synthetic_blocks(5, 6)]
    H.emu_check(oracle, emu7, items, tmp_path, window, seed=window)


def test_emu_v7_handmade_and_fuzz(oracle, emu7, tmp_path):
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
    H.emu_check(oracle, emu7, items, tmp_path, 2048, seed=9)


def test_emu_v7_multi_fragment_and_ragged(oracle, fixtures, emu7, tmp_path):
    """Blocks larger than 64 KiB under one header (offsets cross fragments only through COPY4, which the reference
    never emits but must decode), ragged sizes around the chunk / ring / window sizes, long literals at every alignment."""
    rng = np.random.default_rng(3)
    data = fixtures["corpus/html_x_4"][:300000]
    items = [oracle.compress(data)[1], oracle.compress(fixtures["corpus/alice29.txt"])[1]]
    for n in (1, 14, 15, 16, 60, 61, 62, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097):
        items.append(oracle.compress(rng.integers(0, 256, size=n, dtype=np.uint8).tobytes())[1])
        items.append(oracle.compress((bytes(rng.integers(97, 101, size=7, dtype=np.uint8)) * (n // 7 + 1))[:n])[1])
    # literal runs of every length 500..530 between compressible stretches
    for n in range(500, 531, 3):
        items.append(oracle.compress(b"x" * 100 + rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() + b"y" * 3000)[1])
    H.emu_check(oracle, emu7, items, tmp_path, 2048, seed=11)
