"""Host-side check of the baseline decompress block function (decompress_block_v1: the A/B reference of the fast
engines and their path for blocks above 1 MiB / 2 GiB), compiled with g++ against tests/cpp/simt_emu.h and compared with
the oracle (status, length, bytes, guard bytes around the output).  The GPU parity tests remain the proof for the compiled
kernels; this is what a machine without a GPU can still verify about them."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H
from tests.helpers import BUILD, ROOT, handmade_tag_forms


@pytest.fixture(scope="module")
def emu1():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "emu_v1")
    srcs = [os.path.join(ROOT, "tests", "cpp", "emu_v1.cpp"), os.path.join(ROOT, "tests", "cpp", "simt_emu.h")] + [
        os.path.join(ROOT, "snappier_b200", "csrc", f) for f in ("snp_decompress_v1.cuh", "snp_common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, srcs[0]])
    return exe


def test_emu_v1_blocks(oracle, fixtures, kats, emu1, tmp_path):
    items = [oracle.compress(s)[1] for s in H.edge_strings(kats)]
    items += [oracle.compress(b)[1] for b in (b"", b"a", b"abc" * 100, b"\x00" * 65536)]
    items += H.bad_blocks(oracle, fixtures)
    for name in ("alice29.txt", "html", "kppkn.gtb", "fireworks.jpeg", "geo.protodata"):
        blocks = H.blocks_of(fixtures[f"corpus/{name}"])
        items += [oracle.compress(blocks[0])[1], oracle.compress(blocks[-1])[1]]
    items += [oracle.compress(b)[1] for b in H.synthetic_blocks(5, 6)]
    H.emu_check(oracle, emu1, items, tmp_path, 1, seed=1)


def test_emu_v1_handmade_and_fuzz(oracle, emu1, tmp_path):
    items = handmade_tag_forms()
    rng = np.random.default_rng(8)
    base_blocks = [oracle.compress(b)[1] for b in H.synthetic_blocks(77, 12, size=4096)]
    for i in range(200):
        b = bytearray(base_blocks[i % len(base_blocks)])
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        if i % 5 == 0:
            b = b[: int(rng.integers(0, len(b)))]
        items.append(bytes(b))
    H.emu_check(oracle, emu1, items, tmp_path, 1, seed=9)
