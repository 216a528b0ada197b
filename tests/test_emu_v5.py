"""Host-side check of the DEFAULT decompress engine's logic: the block functions of k_decompress_v5 (sparse-tag prefix
engine + the v3 dense engine) and of the v1 / v3 kernels, compiled with g++ against tests/cpp/simt_emu.h and compared
with the oracle (status, length, bytes, guard bytes around the output).  The GPU parity tests remain the proof for the
compiled kernels; this is what a machine without a GPU can still verify about them."""
from __future__ import annotations

import os
import struct
import subprocess

import numpy as np
import pytest

from tests import helpers as H
from tests.helpers import BUILD, ROOT, handmade_tag_forms


@pytest.fixture(scope="module")
def emu5():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "emu_v5")
    srcs = [os.path.join(ROOT, "tests", "cpp", "emu_v5.cpp"), os.path.join(ROOT, "tests", "cpp", "simt_emu.h")] + [
        os.path.join(ROOT, "snappier_b200", "csrc", f) for f in
        ("snp_decompress_v5.cuh", "snp_decompress_v3.cuh", "snp_decompress_v1.cuh", "snp_common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, srcs[0]])
    return exe


def _run(exe, items, caps, tmp_path, engine, seed):
    rng = np.random.default_rng(seed)
    blob = bytearray(struct.pack("<I", len(items)))
    for b, cap in zip(items, caps):
        blob += struct.pack("<IIII", len(b), cap, int(rng.integers(0, 16)), int(rng.integers(0, 16))) + b
    fin, fout = os.path.join(tmp_path, "batch5.bin"), os.path.join(tmp_path, "result5.bin")
    with open(fin, "wb") as f:
        f.write(blob)
    subprocess.check_call([exe, fin, fout, str(engine)], timeout=1500)
    raw = open(fout, "rb").read()
    res, p = [], 0
    for cap in caps:
        st, n, guard = struct.unpack_from("<iII", raw, p)
        p += 12
        res.append((st, n, guard, raw[p:p + cap]))
        p += cap
    return res


def _check(oracle, exe, items, tmp_path, engine, seed=0):
    caps = []
    for b in items:
        st, n = oracle.uncompressed_length(b)
        caps.append(min(n, 1 << 22) if st == 0 else 0)
    res = _run(exe, items, caps, str(tmp_path), engine, seed)
    for i, (b, cap) in enumerate(zip(items, caps)):
        st, dec = oracle.decompress(b, cap=cap)
        gst, gn, guard, out = res[i]
        assert gst == st, (engine, i, b[:16], gst, st)
        assert guard == 1, (engine, i, "wrote outside its output region")
        assert gn == len(dec) and out[:gn] == dec, (engine, i)


def _bad_blocks(oracle, fixtures):
    bad = [fixtures[f"bad/baddata{i}.snappy"] for i in (1, 2, 3)]
    c = bytearray(oracle.compress(b"making sure we don't crash with corrupted input")[1])
    c[1] -= 1
    c[3] += 1
    bad.append(bytes(c))
    c = bytearray(oracle.compress(b"A" * 1000)[1])
    c[0], c[1] = 255, 127
    bad.append(bytes(c))
    return bad + [b"", b"\x80", b"\xff" * 6, b"\xff\xff\xff\xff\x1f", b"\x05\x10abc", b"\x04\x0cabcd\x01\x00",
                  b"\x08\x0cabcd\x05\x09", b"\x03\x0cabcd", b"\x04\xf0", b"\x0a\x00a\xfe\x01\x00\x00",
                  b"\x40\x00a\xfe\x01\x00", b"\x00garbage", b"\x02\x04ab\x00c"]


@pytest.mark.parametrize("engine", [5, 3, 1])
def test_emu_default_engine_blocks(oracle, fixtures, kats, emu5, tmp_path, engine):
    items = [oracle.compress(s)[1] for s in H.edge_strings(kats)]
    items += [oracle.compress(b)[1] for b in (b"", b"a", b"abc" * 100, b"\x00" * 65536)]
    items += _bad_blocks(oracle, fixtures)
    for name in ("alice29.txt", "html", "kppkn.gtb", "fireworks.jpeg", "geo.protodata"):
        blocks = H.blocks_of(fixtures[f"corpus/{name}"])
        items += [oracle.compress(blocks[0])[1], oracle.compress(blocks[-1])[1]]
    items += [oracle.compress(b)[1] for b in H.synthetic_blocks(5, 6)]
    _check(oracle, emu5, items, tmp_path, engine, seed=engine)


def test_emu_default_engine_handmade_and_fuzz(oracle, emu5, tmp_path):
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
    _check(oracle, emu5, items, tmp_path, 5, seed=9)
