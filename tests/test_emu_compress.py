"""Host-side check of the compress kernels' logic: the fragment functions of k_compress_v3 (default path with the
adaptive batch width, the 16-bit-entry variant) and of the v1 baseline run on
tests/cpp/simt_emu.h; their output must equal the oracle's bytes in both hash modes.  The GPU parity tests remain the
proof for the compiled kernels."""
from __future__ import annotations

import os
import struct
import subprocess

import pytest

from tests import helpers as H
from tests.helpers import BUILD, ROOT


@pytest.fixture(scope="module")
def emuc():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "emu_compress")
    srcs = [os.path.join(ROOT, "tests", "cpp", "emu_compress.cpp"), os.path.join(ROOT, "tests", "cpp", "simt_emu.h")] + [
        os.path.join(ROOT, "snappier_b200", "csrc", f) for f in ("snp_compress_v3.cuh", "snp_compress_v1.cuh", "snp_common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, srcs[0]])
    return exe


def _run(exe, items, tmp_path, variant, hash_mode, w0):
    blob = bytearray(struct.pack("<I", len(items)))
    for b in items:
        blob += struct.pack("<I", len(b)) + b
    fin, fout = os.path.join(tmp_path, "cin.bin"), os.path.join(tmp_path, "cout.bin")
    with open(fin, "wb") as f:
        f.write(blob)
    subprocess.check_call([exe, fin, fout, str(variant), str(hash_mode), str(w0)], timeout=1500)
    raw = open(fout, "rb").read()
    res, p = [], 0
    for _ in items:
        (n,) = struct.unpack_from("<I", raw, p)
        res.append(raw[p + 4:p + 4 + n])
        p += 4 + n
    return res


def _inputs(fixtures, kats):
    items = H.edge_strings(kats)[:7] + [b"", b"a", b"ab" * 7, b"abc" * 100, b"\x00" * 5000]
    items = [b for b in items if len(b) <= 65536]
    for name, lo, n in (("alice29.txt", 1000, 9000), ("html", 0, 12000), ("kppkn.gtb", 500, 7000),
                        ("fireworks.jpeg", 0, 3000), ("geo.protodata", 100, 6000), ("urls.10K", 0, 5000)):
        items.append(fixtures[f"corpus/{name}"][lo:lo + n])
    items += [b[:n] for b, n in zip(H.synthetic_blocks(5, 6, size=8192), (255, 256, 257, 1023, 4097, 8192))]
    return items


@pytest.mark.parametrize("variant,w0", [(3, 16), (3, 32), (3, 1), (6, 16), (1, 32)])
@pytest.mark.parametrize("hash_mode", [0, 1])
def test_emu_compress_fragments_bit_exact(oracle, fixtures, kats, emuc, tmp_path, variant, w0, hash_mode):
    items = _inputs(fixtures, kats)
    got = _run(emuc, items, str(tmp_path), variant, hash_mode, w0)
    for i, b in enumerate(items):
        assert got[i] == oracle.compress(b, hash_mode)[1], (variant, w0, hash_mode, i, len(b))


def test_emu_compress_full_block(oracle, fixtures, emuc, tmp_path):
    """One full 64 KiB block (16384-entry table, probe schedule beyond the first batches) through the default path."""
    b = fixtures["corpus/lcet10.txt"][:65536]
    assert _run(emuc, [b], str(tmp_path), 3, 0, 16)[0] == oracle.compress(b, 0)[1]
