"""Shared test helpers: committed fixtures (tests/golden/) and seeded input generators."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")  # host emulator binaries (git-ignored)
CORPUS = ["alice29.txt", "asyoulik.txt", "fireworks.jpeg", "geo.protodata", "html", "html_x_4",
          "kppkn.gtb", "lcet10.txt", "paper-100k.pdf", "plrabn12.txt", "urls.10K"]


def load_fixtures() -> dict[str, bytes]:
    z = np.load(os.path.join(GOLDEN, "reference_fixtures.npz"))
    return {k: z[k].tobytes() for k in z.files}


def load_kats() -> dict:
    return json.load(open(os.path.join(GOLDEN, "kats.json")))


def load_digests() -> dict:
    return json.load(open(os.path.join(GOLDEN, "oracle_digests.json")))


def parse_framed(stream: bytes):
    """Snappy framing format -> [(chunk_type, body)] (SnappyStreamDecompressor.cs:215-254)."""
    out, i = [], 0
    while i < len(stream):
        t = stream[i]
        n = int.from_bytes(stream[i + 1:i + 4], "little")
        out.append((t, stream[i + 4:i + 4 + n]))
        i += 4 + n
    return out


def golden_blocks(fx: dict[str, bytes]):
    """The 10 raw Snappy blocks inside the two framed goldens -> [(name, masked_crc, block bytes)]."""
    res = []
    for name in ("html_x_4", "alice29"):
        k = 0
        for t, body in parse_framed(fx[f"framed/{name}.snappy"]):
            if t == 0x00:
                res.append((f"{name}[{k}]", int.from_bytes(body[:4], "little"), body[4:]))
                k += 1
    return res


def edge_strings(kats: dict) -> list[bytes]:
    return [(p + f * n + s).encode() for p, f, n, s in kats["edge_strings"]]


def blocks_of(data: bytes, size: int = 65536) -> list[bytes]:
    return [data[i:i + size] for i in range(0, len(data), size)]


def random_data_like_reference(rng: np.random.Generator, n: int) -> bytes:
    """Distribution of SnappyTests.cs:401-446 `RandomData`: runs of a random byte whose
    length is skewed small (the .NET System.Random sequence itself is not reproducible)."""
    out = bytearray()
    while len(out) < n:
        run = 1
        if rng.integers(0, 10) == 0:
            skew = int(rng.integers(0, 9))
            run = int(rng.integers(0, 1 << skew)) + 1
        c = int(rng.integers(0, 256)) if rng.integers(0, 2) else int(rng.integers(0, 4)) + ord("a")
        out += bytes([c]) * run
    return bytes(out[:n])


def synthetic_blocks(seed: int, count: int, size: int = 65536) -> list[bytes]:
    """A mix of block classes (text-like, LZ-synthetic, records, runs, random)."""
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(2, 10)), dtype=np.uint8)) for _ in range(512)]
    res = []
    for i in range(count):
        kind = i % 6
        if kind == 0:  # incompressible
            b = rng.integers(0, 256, size=size, dtype=np.uint8).tobytes()
        elif kind == 1:  # word soup (text-like: many short matches)
            buf = bytearray()
            while len(buf) < size:
                buf += words[int(rng.integers(0, 512))] + b" "
            b = bytes(buf[:size])
        elif kind == 2:  # 64 fresh + 64 copied from 4096 back (BASELINE config 3 shape)
            a = bytearray(rng.integers(0, 256, size=size, dtype=np.uint8).tobytes())
            for s in range(4096, size, 128):
                a[s + 64:s + 128] = a[s + 64 - 4096:s + 128 - 4096]
            b = bytes(a)
        elif kind == 3:  # fixed-width records with slowly varying fields
            rec = np.zeros((size // 32 + 1, 32), np.uint8)
            rec[:, 0:4] = np.arange(rec.shape[0], dtype=np.uint32).view(np.uint8).reshape(-1, 4)
            rec[:, 4:12] = rng.integers(0, 4, size=(rec.shape[0], 8), dtype=np.uint8)
            rec[:, 12:] = 0x20
            b = rec.tobytes()[:size]
        elif kind == 4:  # runs (pattern replication, overlapping copies)
            b = random_data_like_reference(rng, size)
        else:  # short period patterns
            per = int(rng.integers(1, 40))
            pat = rng.integers(0, 256, size=per, dtype=np.uint8).tobytes()
            b = (pat * (size // per + 1))[:size]
        res.append(b)
    return res


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


# ---- host-emulator harness protocol shared by tests/test_emu_v{1,7,8}.py (tests/cpp/emu_v*.cpp) -----------------------

def emu_run(exe, items, caps, tmp_path, engine, seed):
    """Writes the batch file, runs the emulator binary, parses (status, written, guard_ok, output) per item."""
    import struct
    import subprocess
    rng = np.random.default_rng(seed)
    blob = bytearray(struct.pack("<I", len(items)))
    for b, cap in zip(items, caps):
        blob += struct.pack("<IIII", len(b), cap, int(rng.integers(0, 16)), int(rng.integers(0, 16))) + b
    fin, fout = os.path.join(tmp_path, "batch.bin"), os.path.join(tmp_path, "result.bin")
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


def emu_check(oracle, exe, items, tmp_path, engine, seed=0):
    """Status, length, bytes and guard bytes of every item against the oracle."""
    caps = []
    for b in items:
        st, n = oracle.uncompressed_length(b)
        caps.append(min(n, 1 << 22) if st == 0 else 0)
    res = emu_run(exe, items, caps, str(tmp_path), engine, seed)
    for i, (b, cap) in enumerate(zip(items, caps)):
        st, dec = oracle.decompress(b, cap=cap)
        gst, gn, guard, out = res[i]
        assert gst == st, (engine, i, b[:16], gst, st)
        assert guard == 1, (engine, i, "wrote outside its output region")
        assert gn == len(dec) and out[:gn] == dec, (engine, i)


def bad_blocks(oracle, fixtures):
    """The reference's corrupt fixtures and the malformed-stream cases of SnappyTests.cs:212-331."""
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
