"""CPU tests: pin the oracle (oracle/snappy_oracle.c) against every golden vector, KAT and
fixture the reference's own tests hold for the block path."""
import hashlib

import numpy as np
import pytest

from tests import helpers as H


def test_golden_blocks_decompress_and_crc(oracle, fixtures):
    """The 10 blocks inside html_x_4.snappy / alice29.snappy decode to the corpus and every
    masked CRC32C in the chunk headers verifies (SnappyStreamCompressor.cs:233-261)."""
    html4 = fixtures["corpus/html_x_4"]
    alice_crlf = fixtures["corpus/alice29.txt"].replace(b"\n", b"\r\n")
    pos = {"html_x_4": 0, "alice29": 0}
    raw = {"html_x_4": html4, "alice29": alice_crlf}
    blocks = H.golden_blocks(fixtures)
    assert len(blocks) == 10
    for name, crc, blk in blocks:
        key = name.split("[")[0]
        st, dec = oracle.decompress(blk)
        assert st == oracle.OK
        assert dec == raw[key][pos[key]: pos[key] + len(dec)]
        assert oracle.crc32c_masked(dec) == crc
        pos[key] += len(dec)
    assert pos["html_x_4"] == len(html4) and pos["alice29"] == len(alice_crlf)


def test_golden_blocks_compress_mul_hash_byte_exact(oracle, fixtures):
    """MUL-hash CompressFragment reproduces all 10 golden chunks byte for byte -- the only
    compressed bytes the reference pins."""
    for name, _, blk in H.golden_blocks(fixtures):
        st, dec = oracle.decompress(blk)
        st, c = oracle.compress(dec, oracle.HASH_MUL)
        assert st == oracle.OK and c == blk, name


def test_corpus_digests_and_round_trip(oracle, fixtures):
    dig = H.load_digests()
    for f in H.CORPUS:
        d = fixtures["corpus/" + f]
        assert hashlib.sha256(d).hexdigest() == dig[f]["sha256"]
        for mode, key in ((oracle.HASH_CRC32C, "crc32c"), (oracle.HASH_MUL, "mul")):
            st, c = oracle.compress(d, mode)
            assert st == oracle.OK
            assert len(c) == dig[f][key]["len"] and hashlib.sha256(c).hexdigest() == dig[f][key]["sha256"]
            st, back = oracle.decompress(c)
            assert st == oracle.OK and back == d


def test_pyarrow_cross_decompress(oracle, fixtures):
    """Google's C++ Snappy (pyarrow) is an independent decoder of everything the oracle emits,
    and its own output is valid input for the oracle's decoder."""
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")
    for f in H.CORPUS:
        d = fixtures["corpus/" + f]
        for mode in (oracle.HASH_CRC32C, oracle.HASH_MUL):
            st, c = oracle.compress(d, mode)
            assert codec.decompress(c, decompressed_size=len(d)).to_pybytes() == d
        st, back = oracle.decompress(codec.compress(d).to_pybytes())
        assert st == oracle.OK and back == d


def test_find_match_length_kats(oracle, kats):
    """SnappyCompressorTests.cs:10-96."""
    for k in kats["find_match_length"]:
        s1, s2, length = k["s1"].encode(), k["s2"].encode(), k["length"]
        buf = s1 + s2 + b"\0" * max(0, length - len(s2))
        got = oracle.find_match_length(np.frombuffer(buf, np.uint8).copy(), 0, len(s1), len(s1) + length)
        assert got == k["expected"], k


def test_varint_kats(oracle, kats):
    """VarIntEncodingReadTests.cs:7-90, VarIntEncodingWriteTests.cs:5-54."""
    for k in kats["varint"]:
        b = bytes(k["bytes"])
        assert oracle.varint_write(k["value"]) == b
        for pad in (b"", b"\x00" * (16 - len(b)), b"\xff" * (16 - len(b))):
            assert oracle.varint_read(b + pad) == (oracle.OK, k["value"], len(b))
    for b in kats["varint_incomplete"]:
        assert oracle.varint_read(bytes(b))[0] == oracle.INCOMPLETE
    assert oracle.varint_read(b"\xff" * 6)[0] == oracle.INVALID_LENGTH
    assert oracle.varint_read(b"\xff\xff\xff\xff\x1f")[0] == oracle.INVALID_LENGTH
    assert oracle.varint_read(b"")[0] == oracle.INCOMPLETE


def test_crc32c_kats(oracle, kats):
    """Crc32CAlgorithmTests.cs:7-24."""
    for k in kats["crc32c"]:
        assert oracle.crc32c(k["ascii"].encode()) == k["crc"]


def test_hash_table_and_hw_crc_agree(oracle):
    """HashTable.cs:57-71 sizes; the table-driven CRC32C hash equals the SSE4.2 instruction
    the reference executes (HashTable.cs:111) on this host."""
    L = oracle.lib()
    assert [L.orc_table_size(n) for n in (1, 255, 256, 257, 1024, 1025, 16384, 16385, 65536)] == \
        [256, 256, 256, 512, 1024, 2048, 16384, 16384, 16384]
    rng = np.random.default_rng(7)
    for x in rng.integers(0, 2**32, size=5000, dtype=np.uint64):
        for mask in (2 * 255, 2 * 16383):
            assert L.orc_table_hash(int(x), mask, 0) == L.orc_table_hash_fast(int(x), mask, 0)
    assert L.orc_table_hash(0x64636261, 0x7ffe, 1) == ((0x1e35a7bd * 0x64636261 & 0xffffffff) >> 17) & 0x7ffe


def test_tiny_kats_and_edge_strings(oracle, kats):
    """SURVEY App. B tiny KATs + SnappyTests.cs:178-202."""
    want = {b"": "00", b"a": "010061", b"abc": "0308616263",
            b"aaaaaaa" + b"b" * 16 + b"aaaaaabc": "1f0061090100623a01001c6161616161616263"}
    for mode in (oracle.HASH_CRC32C, oracle.HASH_MUL):
        for s, hx in want.items():
            assert oracle.compress(s, mode)[1].hex() == hx
        for s in H.edge_strings(kats):
            st, c = oracle.compress(s, mode)
            assert st == oracle.OK
            assert oracle.decompress(c) == (oracle.OK, s)
    st, c = oracle.compress(b"\0" * 65536)
    assert len(c) == 3077 and hashlib.sha256(c).hexdigest()[:16] == "91f3b2684a367da6"
    st, c = oracle.compress(b"A" * 100000)
    assert len(c) == 4696 and hashlib.sha256(c).hexdigest()[:16] == "4a9e4b62e81e95cb"


def test_bad_data(oracle, fixtures):
    """SnappyTests.cs:212-331."""
    for f in ("baddata1", "baddata2", "baddata3"):
        d = fixtures[f"bad/{f}.snappy"]
        st, n = oracle.uncompressed_length(d)
        assert st == oracle.OK and 0 <= n <= 1 << 20
        assert oracle.decompress(d)[0] == oracle.INVALID_COPY_OFFSET
    # simple corruption (:244-264)
    st, c = oracle.compress(b"making sure we don't crash with corrupted input")
    c = bytearray(c)
    c[1] -= 1
    c[3] += 1
    assert oracle.decompress(bytes(c))[0] in (oracle.INVALID_COPY_OFFSET, oracle.DATA_TOO_LONG, oracle.INCOMPLETE)
    # long length header (:266-285): 16383 declared for a 1000-byte payload
    st, c = oracle.compress(b"A" * 1000)
    c = bytearray(c) + bytes(oracle.get_max_compressed_length(1000) - len(c))
    c[0], c[1] = 255, 127
    assert oracle.decompress(bytes(c), cap=16383)[0] in (oracle.INCOMPLETE, oracle.INVALID_COPY_OFFSET)
    assert oracle.decompress(bytes(c), cap=1000)[0] == oracle.OUTPUT_TOO_SMALL
    # too-small output (:212-242)
    st, c = oracle.compress(b"A" * 100000)
    assert oracle.decompress(c, cap=100)[0] == oracle.OUTPUT_TOO_SMALL
    # zero-length block -> empty output (SnappyDecompressorTests.cs:97-113)
    assert oracle.decompress(b"\x00") == (oracle.OK, b"")
    assert oracle.decompress(b"")[0] == oracle.INCOMPLETE


def test_output_sizing(oracle):
    """SnappyTests.cs:41-118: exact max works, max-5 takes the scratch path, 1024 is too small."""
    rng = np.random.default_rng(3)
    d = rng.integers(0, 256, size=100000, dtype=np.uint8).tobytes()
    full = oracle.get_max_compressed_length(len(d))
    assert oracle.max_compressed_length(65536) == 76491 and oracle.get_max_compressed_length(65536) == 76496
    st, c = oracle.compress(d, cap=full)
    assert st == oracle.OK
    st, c2 = oracle.compress(d, cap=full - 5)
    assert st == oracle.OK and c2 == c
    assert oracle.compress(d, cap=1024) == (oracle.OUTPUT_TOO_SMALL, b"")
    assert oracle.compress(d, cap=len(c))[0] == oracle.OK
    assert oracle.compress(d, cap=len(c) - 1)[0] == oracle.OUTPUT_TOO_SMALL


def test_random_round_trips(oracle):
    """SnappyTests.cs:401-446 (distribution re-created, see helpers.random_data_like_reference)."""
    rng = np.random.default_rng(301)
    for i in range(300):
        n = int(rng.integers(65536, 131072)) if i < 20 else int(rng.integers(0, 4096))
        d = H.random_data_like_reference(rng, n)
        for mode in (oracle.HASH_CRC32C, oracle.HASH_MUL):
            st, c = oracle.compress(d, mode)
            assert st == oracle.OK and oracle.decompress(c) == (oracle.OK, d)


def test_batch_driver_matches_single_calls(oracle):
    blocks = H.synthetic_blocks(11, 24)
    from snappier_b200.batch import pack
    base, off, ln = pack(blocks)
    caps = np.full(len(blocks), oracle.get_max_compressed_length(65536), np.uint32)
    out_off = np.arange(len(blocks), dtype=np.uint64) * caps[0]
    out = np.zeros(int(caps.sum()), np.uint8)
    bad, out_len, status = oracle.compress_batch(base, off, ln, out, out_off, caps, threads=4)
    assert bad == 0 and not status.any()
    comp = [out[int(o):int(o) + int(l)].tobytes() for o, l in zip(out_off, out_len)]
    assert comp == [oracle.compress(b)[1] for b in blocks]
    cbase, coff, clen = pack(comp)
    dout = np.zeros(65536 * len(blocks), np.uint8)
    bad, dlen, status = oracle.decompress_batch(cbase, coff, clen, dout, np.arange(len(blocks), dtype=np.uint64) * 65536,
                                                np.full(len(blocks), 65536, np.uint32), threads=3)
    assert bad == 0 and dout.tobytes() == b"".join(blocks)


def test_framing_checker_reproduces_golden_streams(oracle, fixtures):
    """The oracle-side framing checker (pyoracle.frame_*) is pinned by the reference's two framed
    goldens: MUL-hash frame_compress reproduces html_x_4.snappy and alice29.snappy byte for byte
    (stream identifier, chunk headers, masked CRCs, payloads), and frame_decompress inverts them."""
    html4 = fixtures["corpus/html_x_4"]
    alice_crlf = fixtures["corpus/alice29.txt"].replace(b"\n", b"\r\n")
    for raw, name in ((html4, "html_x_4"), (alice_crlf, "alice29")):
        gold = fixtures[f"framed/{name}.snappy"]
        assert oracle.frame_compress(raw, oracle.HASH_MUL) == gold
        assert oracle.frame_decompress(gold) == (oracle.OK, raw)
        assert oracle.frame_decompress(oracle.frame_compress(raw, oracle.HASH_CRC32C)) == (oracle.OK, raw)
    # incompressible chunks fall back to type 0x01 with 8 + len bytes (SnappyStreamCompressorTests.cs:7-46)
    rnd = np.random.default_rng(1).integers(0, 256, size=256, dtype=np.uint8).tobytes()
    assert len(oracle.frame_compress(rnd)) == 10 + 8 + 256
