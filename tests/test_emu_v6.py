"""Host-side check of the v6 decompress kernels' LOGIC (snappier_b200/csrc/snp_decompress_v6.cuh).

The warp-level device functions are compiled with g++ against tests/cpp/simt_emu.h (32 lanes as
coroutines, collectives as rendezvous) and run over real Snappy blocks; status, length, bytes and
the guard bytes around every output region are compared with the oracle.  This is what lets the
parse / dependency-round / window / flush logic be debugged without a GPU; the GPU parity tests
(tests/test_gpu_parity.py) remain the proof for the compiled kernel.
"""
from __future__ import annotations

import os
import struct
import subprocess

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")


def _build(name: str, extra: list[str]) -> str:
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    srcs = [os.path.join(ROOT, "tests", "cpp", "emu_v6.cpp"), os.path.join(ROOT, "tests", "cpp", "simt_emu.h"),
            os.path.join(ROOT, "snappier_b200", "csrc", "snp_decompress_v6.cuh"),
            os.path.join(ROOT, "snappier_b200", "csrc", "snp_common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, srcs[0]] + extra)
    return exe


def run_emu(exe: str, items: list[bytes], caps: list[int], tmp_path, seed: int = 0):
    rng = np.random.default_rng(seed)
    blob = bytearray(struct.pack("<I", len(items)))
    for b, cap in zip(items, caps):
        blob += struct.pack("<IIII", len(b), cap, int(rng.integers(0, 16)), int(rng.integers(0, 16))) + b
    fin, fout = os.path.join(tmp_path, "batch.bin"), os.path.join(tmp_path, "result.bin")
    with open(fin, "wb") as f:
        f.write(blob)
    subprocess.check_call([exe, fin, fout], timeout=1500)
    raw = open(fout, "rb").read()
    res, p = [], 0
    for cap in caps:
        st, n, guard = struct.unpack_from("<iII", raw, p)
        p += 12
        res.append((st, n, guard, raw[p:p + cap]))
        p += cap
    return res


def check(oracle, exe, items, tmp_path, caps=None, seed=0, allow_fallback=False):
    if caps is None:
        caps = []
        for b in items:
            st, n = oracle.uncompressed_length(b)
            caps.append(min(n, 1 << 22) if st == 0 else 0)
    res = run_emu(exe, items, caps, str(tmp_path), seed)
    n_fb = 0
    for i, (b, cap) in enumerate(zip(items, caps)):
        st, dec = oracle.decompress(b, cap=cap)
        gst, gn, guard, out = res[i]
        if allow_fallback and gst == 100:
            n_fb += 1
            continue
        assert gst == st, (i, b[:16], gst, st)
        assert guard == 1, (i, "wrote outside its output region")
        assert gn == len(dec), (i, gn, len(dec))
        assert out[:gn] == dec, (i, "bytes differ")
    return n_fb


@pytest.fixture(scope="module")
def emu():
    return _build("emu_v6", [])


def test_emu_v6_edge_and_bad_blocks(oracle, fixtures, kats, emu, tmp_path):
    items = [oracle.compress(s)[1] for s in H.edge_strings(kats)]
    items += [oracle.compress(b)[1] for b in (b"", b"a", b"ab" * 7, b"abc" * 100, b"\x00" * 65536, b"xyz" * 30000)]
    bad = [fixtures[f"bad/baddata{i}.snappy"] for i in (1, 2, 3)]
    c = bytearray(oracle.compress(b"making sure we don't crash with corrupted input")[1])
    c[1] -= 1
    c[3] += 1
    bad.append(bytes(c))
    c = bytearray(oracle.compress(b"A" * 1000)[1])
    c[0], c[1] = 255, 127
    bad.append(bytes(c))
    bad += [b"", b"\x80", b"\xff" * 6, b"\xff\xff\xff\xff\x1f", b"\x05\x10abc", b"\x04\x0cabcd\x01\x00",
            b"\x08\x0cabcd\x05\x09", b"\x03\x0cabcd", b"\x04\xf0", b"\x0a\x00a\xfe\x01\x00\x00",
            b"\x40\x00a\xfe\x01\x00", b"\x00garbage", b"\x02\x04ab\x00c"]
    good = oracle.compress(b"interleaved good block " * 100)[1]
    for b in bad:
        items += [b, good]
    check(oracle, emu, items, tmp_path)


def test_emu_v6_corpus_and_synthetic_blocks(oracle, fixtures, emu, tmp_path):
    items = []
    for name in H.CORPUS:
        blocks = H.blocks_of(fixtures[f"corpus/{name}"])
        for b in (blocks[0], blocks[-1]):
            items.append(oracle.compress(b)[1])
    items += [oracle.compress(b)[1] for b in H.synthetic_blocks(5, 12)]
    items += [oracle.compress(b, oracle.HASH_MUL)[1] for b in H.synthetic_blocks(6, 6, size=20000)]
    check(oracle, emu, items, tmp_path, seed=1)


def handmade_tag_forms() -> list[bytes]:
    """Hand-assembled blocks: COPY4, multi-byte literal lengths, literals around the 64 / 128-byte path boundaries,
    offsets around 16, literals > 64 bytes at every slot position, chained near / far copies."""
    rng = np.random.default_rng(12)
    lit = rng.integers(0, 256, size=70000, dtype=np.uint8).tobytes()

    def varint(v):
        out = bytearray()
        while v >= 0x80:
            out.append((v & 0x7f) | 0x80)
            v >>= 7
        out.append(v)
        return bytes(out)

    def literal(data):
        n = len(data) - 1
        if n < 60:
            return bytes([n << 2]) + data
        k = (n.bit_length() + 7) // 8
        return bytes([(59 + k) << 2]) + n.to_bytes(k, "little") + data

    def copy(off, ln):
        return bytes([((ln - 1) << 2) | 2]) + off.to_bytes(2, "little")

    items = []
    body = bytes([62 << 2]) + (len(lit) - 1).to_bytes(3, "little") + lit
    body += bytes([((10 - 1) << 2) | 3]) + (69000).to_bytes(4, "little")
    body += bytes([((64 - 1) << 2) | 3]) + (3).to_bytes(4, "little")
    items.append(varint(70000 + 10 + 64) + body)
    items.append(varint(70000) + bytes([63 << 2]) + (len(lit) - 1).to_bytes(4, "little") + lit)
    items.append(varint(300) + bytes([61 << 2]) + (299).to_bytes(2, "little") + lit[:300])
    items.append(varint(20) + bytes([3 << 2]) + b"abcd" + bytes([(4 - 1) << 2 | 3]) + (0).to_bytes(4, "little"))
    items.append(varint(20) + bytes([3 << 2]) + b"abcd" + bytes([(4 - 1) << 2 | 3]) + (5).to_bytes(4, "little"))
    # every copy offset 1..40 x lengths around the 16-byte trips, each after literals of boundary sizes
    for lit_len in (1, 15, 16, 17, 63, 64, 65, 127, 128, 129, 200, 1100, 3000):
        body, total = bytearray(), 0
        body += literal(lit[:lit_len])
        total += lit_len
        for off in list(range(1, 41)) + [63, 64, 65, 100]:
            for ln in (1, 4, 15, 16, 17, 31, 32, 33, 48, 63, 64):
                if off <= total:
                    body += copy(off, ln)
                    total += ln
            body += literal(lit[total % 5000: total % 5000 + (off % 7) + 1])
            total += (off % 7) + 1
        items.append(varint(total) + bytes(body))
    # a literal > 64 bytes at every slot position around the group boundary (head + length never split: pad slot),
    # with lengths whose low byte looks like a literal head (0x80) or is zero
    for pre in range(27, 36):
        for big in (65, 128, 256, 0x180, 1000):
            body, total = bytearray(), 0
            for i in range(pre):
                body += literal(lit[i:i + 1 + (i % 3)])
                total += 1 + (i % 3)
            body += literal(lit[100:100 + big])
            total += big
            for off, ln in ((1, 20), (big, 33), (5, 4), (total // 2, 64)):
                body += copy(off, ln)
                total += ln
            body += literal(lit[7:7 + big + 3])
            total += big + 3
            items.append(varint(total) + bytes(body))
    # long runs of chained far/near copies and literals > 64 interleaved (window slide + re-seed)
    body, total = bytearray(), 0
    for i in range(400):
        n = int(rng.integers(1, 300))
        body += literal(lit[i * 100: i * 100 + n])
        total += n
        for _ in range(int(rng.integers(0, 6))):
            off = int(rng.integers(1, min(total, 65535) + 1))
            ln = int(rng.integers(1, 65))
            body += copy(off, ln)
            total += ln
    items.append(varint(total) + bytes(body))
    return items



def test_emu_v6_handmade_tag_forms(oracle, emu, tmp_path):
    check(oracle, emu, handmade_tag_forms(), tmp_path, seed=2)


def test_emu_v6_fuzz(oracle, emu, tmp_path):
    rng = np.random.default_rng(99)
    items = []
    base_blocks = [oracle.compress(b)[1] for b in H.synthetic_blocks(77, 12, size=4096)]
    for i in range(300):
        b = bytearray(base_blocks[i % len(base_blocks)])
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        if i % 5 == 0:
            b = b[: int(rng.integers(0, len(b)))]
        items.append(bytes(b))
    for i in range(100):
        n = int(rng.integers(1, 64))
        items.append(bytes([int(rng.integers(1, 40))]) + rng.integers(0, 256, size=n, dtype=np.uint8).tobytes())
    caps = []
    for b in items:
        st, n = oracle.uncompressed_length(b)
        caps.append(min(n, 1 << 16) if st == 0 else 0)
    check(oracle, emu, items, tmp_path, caps=caps, seed=3)


def test_emu_v6_checkpoint_budget_fallback(oracle, tmp_path):
    """With a 4-group budget every block above 128 tags must be handed to the fallback engine, the rest decode."""
    exe = _build("emu_v6_ckb4", ["-DSNP6_CKB=4u"])
    items = [oracle.compress(b)[1] for b in H.synthetic_blocks(9, 6, size=3000)]
    items += [oracle.compress(b"tiny"), oracle.compress(b"abcdefgh" * 10)]
    items = [b if isinstance(b, bytes) else b[1] for b in items]
    n_fb = check(oracle, exe, items, tmp_path, seed=4, allow_fallback=True)
    assert 0 < n_fb < len(items)
