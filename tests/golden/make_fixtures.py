#!/usr/bin/env python3
"""Regenerate tests/golden/* from the reference checkout (run in the CPU container only).

    python tests/golden/make_fixtures.py [/root/reference]

/root/reference does not exist on the GPU box, so everything the tests need from
it is frozen here:

  reference_fixtures.npz   the reference's own test data, byte-for-byte, packed
                           into one deflate-compressed archive:
                             framed/html_x_4.snappy, framed/alice29.snappy
                                 (golden framed streams; Snappier.Benchmarks/DecompressHtml.cs:19)
                             bad/baddata{1,2,3}.snappy   (Snappier.Tests/SnappyTests.cs:287-331)
                             corpus/<11 files>           (Snappier.Tests/SnappyTests.cs:8-20)
  kats.json                known-answer vectors lifted from the reference's unit
                           tests by regex (FindMatchLength x45, varint x13 +
                           incomplete x10, CRC32C x4, edge strings x9)
  oracle_digests.json      lengths + sha256 of the ORACLE's compressed output per
                           corpus file and hash mode (regression pins; the MUL
                           column is anchored by the framed goldens, the CRC32C
                           column is "parity unpinned" -- see oracle/snappy_oracle.h)
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CORPUS = ["alice29.txt", "asyoulik.txt", "fireworks.jpeg", "geo.protodata", "html", "html_x_4",
          "kppkn.gtb", "lcet10.txt", "paper-100k.pdf", "plrabn12.txt", "urls.10K"]


def main(ref: str) -> None:
    td = os.path.join(ref, "Snappier.Tests", "TestData")
    arrays = {}
    for f in ("html_x_4.snappy", "alice29.snappy"):
        arrays["framed/" + f] = np.fromfile(os.path.join(td, f), np.uint8)
    for f in ("baddata1.snappy", "baddata2.snappy", "baddata3.snappy"):
        arrays["bad/" + f] = np.fromfile(os.path.join(td, f), np.uint8)
    for f in CORPUS:
        arrays["corpus/" + f] = np.fromfile(os.path.join(td, f), np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_fixtures.npz"), **arrays)

    kats: dict = {}
    # FindMatchLength: [InlineData(6, "012345", "012345", 6)]
    src = open(os.path.join(ref, "Snappier.Tests/Internal/SnappyCompressorTests.cs"), encoding="utf-8-sig").read()
    kats["find_match_length"] = [
        {"expected": int(m[1]), "s1": m[2], "s2": m[3], "length": int(m[4])}
        for m in re.finditer(r'\[InlineData\((\d+), "([^"]*)", "([^"]*)", (\d+)\)\]', src)]
    assert len(kats["find_match_length"]) == 45, len(kats["find_match_length"])
    # varint: { 0x555, [ 0xD5, 0x0A ] },
    src = open(os.path.join(ref, "Snappier.Tests/Internal/VarIntEncodingReadTests.cs"), encoding="utf-8-sig").read()
    body = src[src.index("TestData()"): src.index("IncompleteTestData()")]
    kats["varint"] = [
        {"value": int(m[1], 16), "bytes": [int(x, 16) for x in re.findall(r"0[xX][0-9A-Fa-f]+", m[2])]}
        for m in re.finditer(r"\{\s*(0x[0-9A-Fa-f]+),\s*\[([^\]]*)\]\s*\}", body)]
    assert len(kats["varint"]) == 13
    body = src[src.index("IncompleteTestData()"): src.index("Test_TryRead(")]
    kats["varint_incomplete"] = [
        [int(x, 16) for x in re.findall(r"0[xX][0-9A-Fa-f]+", m[1])]
        for m in re.finditer(r"\{\s*\[([^\]]*)\]\s*\}", body)]
    assert len(kats["varint_incomplete"]) == 10
    # CRC32C: [InlineData("123456789", 0xe3069283)]
    src = open(os.path.join(ref, "Snappier.Tests/Internal/Crc32CAlgorithmTests.cs"), encoding="utf-8-sig").read()
    kats["crc32c"] = [{"ascii": m[1], "crc": int(m[2], 16)}
                      for m in re.finditer(r'\[InlineData\("([^"]*)", (0x[0-9a-fA-F]+)\)\]', src)]
    assert len(kats["crc32c"]) == 4
    # Edge strings (SnappyTests.cs:178-189), as (prefix, fill char, fill count, suffix)
    kats["edge_strings"] = [
        ["", "", 0, ""], ["a", "", 0, ""], ["ab", "", 0, ""], ["abc", "", 0, ""],
        ["aaaaaaa", "b", 16, "aaaaaabc"], ["aaaaaaa", "b", 256, "aaaaaabc"],
        ["aaaaaaa", "b", 2047, "aaaaaabc"], ["aaaaaaa", "b", 65536, "aaaaaabc"],
        ["abcaaaaaaa", "b", 65536, "aaaaaabc"]]
    json.dump(kats, open(os.path.join(HERE, "kats.json"), "w"), indent=1)

    from oracle import pyoracle as O
    O.build()
    dig = {}
    for f in CORPUS:
        d = arrays["corpus/" + f].tobytes()
        e = {"len": len(d), "sha256": hashlib.sha256(d).hexdigest()}
        for mode, key in ((O.HASH_CRC32C, "crc32c"), (O.HASH_MUL, "mul")):
            st, c = O.compress(d, mode)
            assert st == 0
            frags = []
            for i in range(0, len(d), 65536):
                st, cf = O.compress(d[i:i + 65536], mode)
                frags.append([len(cf), hashlib.sha256(cf).hexdigest()[:16]])
            e[key] = {"len": len(c), "sha256": hashlib.sha256(c).hexdigest(), "blocks": frags}
        dig[f] = e
    json.dump(dig, open(os.path.join(HERE, "oracle_digests.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
