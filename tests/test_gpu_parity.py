"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle on
the same inputs -- bit-exact for compressed bytes (both hash modes), decompressed bytes,
lengths and per-block status.  Nothing here reads /root/reference."""
import hashlib

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _corpus_blocks(fixtures):
    blocks, names = [], []
    for f in H.CORPUS:
        for i, b in enumerate(H.blocks_of(fixtures["corpus/" + f])):
            blocks.append(b)
            names.append(f"{f}[{i}]")
    return names, blocks


@pytest.mark.parametrize("mode", [0, 1], ids=["crc32c", "mul"])
def test_compress_corpus_blocks_bit_exact(engine, oracle, fixtures, mode):
    """Every 64 KiB fragment of the reference's 11 corpus files (SnappyTests.cs:8-20)."""
    from snappier_b200.batch import compress_many
    names, blocks = _corpus_blocks(fixtures)
    got, status = compress_many(engine, blocks, mode)
    assert not status.any()
    for name, b, g in zip(names, blocks, got):
        st, want = oracle.compress(b, mode)
        assert g == want, name


def test_compress_golden_chunks_mul_hash(engine, oracle, fixtures):
    """MUL-hash output equals the reference's own golden chunks byte for byte."""
    from snappier_b200.batch import compress_many, decompress_many
    gold = H.golden_blocks(fixtures)
    raw, status = decompress_many(engine, [b for _, _, b in gold])
    assert not status.any()
    for (name, crc, blk), r in zip(gold, raw):
        assert oracle.crc32c_masked(r) == crc, name
    comp, status = compress_many(engine, raw, 1)
    assert not status.any()
    assert comp == [b for _, _, b in gold]


def test_decompress_corpus_blocks(engine, oracle, fixtures):
    from snappier_b200.batch import decompress_many
    names, blocks = _corpus_blocks(fixtures)
    for mode in (0, 1):
        comp = [oracle.compress(b, mode)[1] for b in blocks]
        got, status = decompress_many(engine, comp)
        assert not status.any()
        assert got == blocks
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")  # blocks produced by Google's encoder (different tag mix)
    comp = [codec.compress(b).to_pybytes() for b in blocks]
    got, status = decompress_many(engine, comp)
    assert not status.any() and got == blocks


@pytest.mark.parametrize("mode", [0, 1], ids=["crc32c", "mul"])
def test_synthetic_blocks_round_trip_and_parity(engine, oracle, mode):
    from snappier_b200.batch import compress_many, decompress_many
    blocks = H.synthetic_blocks(1234 + mode, 96)
    # ragged sizes, including the < 15-byte and table-size boundaries (HashTable.cs:57-71)
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 3, 4, 14, 15, 16, 17, 31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 4095, 4096, 4097,
              16383, 16384, 16385, 65521, 65535):
        blocks.append(blocks[int(rng.integers(0, 96))][:n])
    got, status = compress_many(engine, blocks, mode)
    assert not status.any()
    want = [oracle.compress(b, mode)[1] for b in blocks]
    bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, bad[:10]
    back, status = decompress_many(engine, got)
    assert not status.any() and back == blocks


def test_edge_strings_single_call_api(oracle, kats):
    """SnappyTests.cs:178-202 through the Snappy facade mirror (multi-fragment inputs included)."""
    from snappier_b200 import snappy as S
    for s in H.edge_strings(kats):
        for mode in (0, 1):
            c = S.compress_to_array(s, mode)
            assert c == oracle.compress(s, mode)[1]
            assert S.decompress_to_array(c) == s
            assert S.decompress_to_memory(c).tobytes() == s


def test_whole_files_single_call_api(oracle, fixtures):
    """Snappy.Compress of inputs > 64 KiB = varint ++ independent fragments (SnappyCompressor.cs:40-80);
    digests pinned in tests/golden/oracle_digests.json."""
    from snappier_b200 import snappy as S
    dig = H.load_digests()
    for f in H.CORPUS:
        d = fixtures["corpus/" + f]
        for mode, key in ((0, "crc32c"), (1, "mul")):
            c = S.compress_to_array(d, mode)
            assert len(c) == dig[f][key]["len"] and hashlib.sha256(c).hexdigest() == dig[f][key]["sha256"], (f, key)
        assert S.decompress_to_array(c) == d


def _sequence_fragments(seg_lens):
    """The fragment partition of SnappyCompressor.Compress(ReadOnlySequence<byte>, ..), restated from
    SnappyCompressor.cs:103-143: the next fragment is the first segment's part of the next <= 64 KiB when that part is the
    whole fragment (fragment.IsSingleSegment) or at least 32 KiB, otherwise the whole (copied) fragment."""
    segs = [n for n in seg_lens if n]
    frags, i, o, left = [], 0, 0, sum(segs)
    while left:
        frag = min(left, 65536)
        first = min(segs[i] - o, frag)
        take = first if (first == frag or first >= 32768) else frag
        frags.append(take)
        left -= take
        while take:
            step = min(segs[i] - o, take)
            o += step
            take -= step
            if o == segs[i]:
                i, o = i + 1, 0
    return frags


def test_sequence_overloads_follow_the_segmentation(oracle, fixtures):
    """Snappy.Compress(ReadOnlySequence, IBufferWriter) / DecompressToMemory(ReadOnlySequence) (SnappyTests.cs:122-174 and
    :333-399 split the input into 16/32/64 KiB and 1024-byte segments): bytes equal varint ++ the oracle's
    CompressFragment of every fragment of the reference's partition, for regular and ragged segmentations."""
    from snappier_b200 import snappy as S
    data = fixtures["corpus/lcet10.txt"] + fixtures["corpus/kppkn.gtb"][:100000]
    rng = np.random.default_rng(31)
    plans = [[len(data)], [16384] * 40, [32768] * 20, [65536] * 10, [1024] * 700, [40000, 1000, 70000, 5, 200000],
             [32767, 32769, 65536, 1, 65535, 131072], [int(x) for x in rng.integers(1, 90000, size=64)]]
    for plan in plans:
        segs, o = [], 0
        for n in plan:
            if o >= len(data):
                break
            segs.append(data[o:o + n])
            o += n
        if o < len(data):
            segs.append(data[o:])
        want = bytearray(oracle.varint_write(len(data)))
        o = 0
        for n in _sequence_fragments([len(x) for x in segs]):
            buf = np.zeros(oracle.max_compressed_length(n), np.uint8)
            frag = np.frombuffer(data[o:o + n], np.uint8)
            w = oracle.lib().orc_compress_fragment(frag.ctypes.data, n, buf.ctypes.data, 0)
            want += buf[:w].tobytes()
            o += n
        got = S.compress_sequence(segs)
        assert got == bytes(want), plan[:6]
        assert oracle.decompress(got)[1] == data
        # the block arrives split at arbitrary points (SnappyTests.cs:357-378: 1024-byte segments)
        cuts = sorted(set(int(x) for x in rng.integers(0, len(got), size=9)) | {0, 1, 2, len(got)})
        pieces = [got[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
        assert S.decompress_sequence(pieces) == data
    assert S.compress_sequence([data]) == S.compress_to_array(data)  # one segment == the span overload
    assert S.compress_sequence([]) == b"\x00" and S.decompress_sequence([b"\x00"]) == b""
    with pytest.raises(S.InvalidDataException):
        S.decompress_sequence([got[:100], got[100:200]])  # truncated block


def test_output_sizing_semantics(oracle):
    """SnappyTests.cs:41-118."""
    from snappier_b200 import snappy as S
    rng = np.random.default_rng(3)
    d = rng.integers(0, 256, size=100000, dtype=np.uint8).tobytes()
    full = S.get_max_compressed_length(len(d))
    out = np.zeros(full, np.uint8)
    n = S.compress(d, out)
    want = oracle.compress(d)[1]
    assert out[:n].tobytes() == want
    out2 = np.zeros(full - 5, np.uint8)
    assert S.compress(d, out2) == n and out2[:n].tobytes() == want
    with pytest.raises(S.ArgumentException):
        S.compress(d, np.zeros(1024, np.uint8))
    assert S.try_compress(d, np.zeros(1024, np.uint8)) == (False, 0)
    assert S.try_compress(d, np.zeros(0, np.uint8)) == (False, 0)
    assert S.try_compress(d, np.zeros(n, np.uint8))[0] is True
    assert S.try_compress(d, np.zeros(n - 1, np.uint8)) == (False, 0)
    # single-fragment bounded slot
    small = d[:5000]
    k = len(oracle.compress(small)[1])
    assert S.try_compress(small, np.zeros(k, np.uint8)) == (True, k)
    assert S.try_compress(small, np.zeros(k - 1, np.uint8)) == (False, 0)
    # overlap (SnappyTests.cs:204-210)
    buf = np.zeros(1024, np.uint8)
    with pytest.raises(S.InvalidOperationException):
        S.compress(buf, buf[1023:])


def test_bad_data_statuses(engine, oracle, fixtures):
    """SnappyTests.cs:212-331: same status (hence exception type) as the oracle for every bad input."""
    from snappier_b200 import snappy as S
    from snappier_b200.batch import decompress_many
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
    items = []
    for b in bad:
        items += [b, good]
    caps = []
    for b in items:
        st, n = oracle.uncompressed_length(b)
        caps.append(min(n, 1 << 20) if st == 0 else 0)
    got, status = decompress_many(engine, items, caps)
    for i, b in enumerate(items):
        st, dec = oracle.decompress(b, cap=caps[i])
        assert status[i] == st, (i, b[:16], status[i], st)
        assert got[i] == dec
    # a corrupt item never poisons its neighbours
    assert all(got[i] == b"interleaved good block " * 100 for i in range(1, len(items), 2))
    for b in bad[:4]:
        with pytest.raises(S.InvalidDataException):
            S.decompress_to_array(b)
    with pytest.raises(S.InvalidDataException):
        S.decompress(bad[4], np.zeros(16383, np.uint8))
    # too-small output: ArgumentException / TryDecompress false (SnappyTests.cs:212-242)
    c = S.compress_to_array(b"A" * 100000)
    with pytest.raises(S.ArgumentException):
        S.decompress(c, np.zeros(100, np.uint8))
    out = np.zeros(100, np.uint8)
    assert S.try_decompress(c, out) == (False, 100) and out.tobytes() == b"A" * 100
    assert S.decompress_to_array(b"\x00") == b""


def test_copy4_long_literals_and_big_blocks(engine, oracle):
    """Tag forms the reference compressor never emits but its decoder accepts (SnappyDecompressor.cs:305-313:
    COPY4; :278-288: 2/3/4-byte literal lengths), plus blocks far above 64 KiB decoded by one warp."""
    from snappier_b200.batch import decompress_many
    rng = np.random.default_rng(12)
    lit = rng.integers(0, 256, size=70000, dtype=np.uint8).tobytes()

    def varint(v):
        out = bytearray()
        while v >= 0x80:
            out.append((v & 0x7f) | 0x80)
            v >>= 7
        out.append(v)
        return bytes(out)

    items = []
    # literal with a 3-byte length, then COPY4 reaching 69000 bytes back, then an overlapping COPY4 (pattern fill)
    body = bytes([62 << 2]) + (len(lit) - 1).to_bytes(3, "little") + lit
    body += bytes([((10 - 1) << 2) | 3]) + (69000).to_bytes(4, "little")
    body += bytes([((64 - 1) << 2) | 3]) + (3).to_bytes(4, "little")
    items.append(varint(70000 + 10 + 64) + body)
    # literal with a 4-byte length field; literal with a 2-byte length field
    items.append(varint(70000) + bytes([63 << 2]) + (len(lit) - 1).to_bytes(4, "little") + lit)
    items.append(varint(300) + bytes([61 << 2]) + (299).to_bytes(2, "little") + lit[:300])
    # COPY4 with offset 0 / beyond the produced data -> invalid copy offset
    items.append(varint(20) + bytes([3 << 2]) + b"abcd" + bytes([(4 - 1) << 2 | 3]) + (0).to_bytes(4, "little"))
    items.append(varint(20) + bytes([3 << 2]) + b"abcd" + bytes([(4 - 1) << 2 | 3]) + (5).to_bytes(4, "little"))
    # multi-megabyte blocks: repetitive (long copies) and text-like
    items.append(oracle.compress(b"0123456789abcdef" * 200000)[1])
    items.append(oracle.compress(b"".join(H.synthetic_blocks(3, 40)))[1])
    caps = [oracle.uncompressed_length(b)[1] for b in items]
    got, status = decompress_many(engine, items, caps)
    for i, b in enumerate(items):
        st, dec = oracle.decompress(b, cap=caps[i])
        assert status[i] == st and got[i] == dec, i
    assert status[0] == 0 and status[3] == oracle.INVALID_COPY_OFFSET and status[4] == oracle.INVALID_COPY_OFFSET


def test_hypothesis_fuzz_decoder_never_diverges(engine, oracle):
    """Random byte soup and mutated valid blocks: status and output always equal the oracle's."""
    from snappier_b200.batch import decompress_many
    rng = np.random.default_rng(99)
    items = []
    base_blocks = [oracle.compress(b)[1] for b in H.synthetic_blocks(77, 12, size=4096)]
    for i in range(600):
        b = bytearray(base_blocks[i % len(base_blocks)])
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        if i % 5 == 0:
            b = b[: int(rng.integers(0, len(b)))]
        items.append(bytes(b))
    for i in range(200):
        n = int(rng.integers(1, 64))
        items.append(bytes([int(rng.integers(1, 40))]) + rng.integers(0, 256, size=n, dtype=np.uint8).tobytes())
    caps = []
    for b in items:
        st, n = oracle.uncompressed_length(b)
        caps.append(min(n, 1 << 16) if st == 0 else 0)
    got, status = decompress_many(engine, items, caps)
    for i, b in enumerate(items):
        st, dec = oracle.decompress(b, cap=caps[i])
        assert status[i] == st, (i, status[i], st)
        assert got[i] == dec, i


def test_device_mode_large_batch_properties(engine, oracle):
    """Device-resident batch (the throughput path): compress -> decompress round trip over 4096
    blocks, per-block parity on a sample, and a checksum of checksums over the full batch."""
    import torch
    blocks = H.synthetic_blocks(2024, 64)
    n = 4096
    dev = torch.device("cuda:0")
    src = torch.from_numpy(np.frombuffer(b"".join(blocks), np.uint8).copy()).to(dev)
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    in_off = (idx % len(blocks)) * 65536
    in_len = torch.full((n,), 65536, dtype=torch.int32, device=dev)
    pitch = 76496
    comp = torch.zeros(n * pitch, dtype=torch.uint8, device=dev)
    c_off = idx * pitch
    c_cap = torch.full((n,), pitch, dtype=torch.int32, device=dev)
    c_len = torch.zeros(n, dtype=torch.int32, device=dev)
    status = torch.full((n,), -99, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    before = engine.launch_count
    engine.compress_batch_device(src, in_off, in_len, comp, c_off, c_cap, c_len, status, 0, stream)
    out = torch.zeros(n * 65536, dtype=torch.uint8, device=dev)
    o_off = idx * 65536
    o_cap = torch.full((n,), 65536, dtype=torch.int32, device=dev)
    o_len = torch.zeros(n, dtype=torch.int32, device=dev)
    status2 = torch.full((n,), -99, dtype=torch.int32, device=dev)
    engine.decompress_batch_device(comp, c_off, c_len, out, o_off, o_cap, o_len, status2, stream)
    torch.cuda.synchronize()
    assert engine.launch_count - before == 2
    assert int(status.abs().sum()) == 0 and int(status2.abs().sum()) == 0
    assert bool((o_len == 65536).all())
    want_len = np.array([len(oracle.compress(b)[1]) for b in blocks])
    assert np.array_equal(c_len.cpu().numpy(), want_len[np.arange(n) % len(blocks)])
    assert torch.equal(out.view(n // len(blocks), -1), src.view(1, -1).expand(n // len(blocks), -1))
    cl = c_len.cpu().numpy()
    ch = comp.view(n, pitch)
    for i in (0, 1, 63, 64, 2047, 4095):
        assert ch[i, : cl[i]].cpu().numpy().tobytes() == oracle.compress(blocks[i % len(blocks)])[1]
    # the 2 KiB-window instantiation of the default kernel and the lane-per-block engine on the same device-resident batch
    for env in ({"SNP_V7_WINDOW": "2048"}, {"SNP_DECOMP_KERNEL": "8"}, {"SNP_DECOMP_KERNEL": "8", "SNP_V8_CFG": "2"}):
        ex = _engine_with(env)
        out.zero_()
        o_len.zero_()
        status2.fill_(-99)
        ex.decompress_batch_device(comp, c_off, c_len, out, o_off, o_cap, o_len, status2, stream)
        torch.cuda.synchronize()
        assert ex.launch_count == 1
        assert int(status2.abs().sum()) == 0 and bool((o_len == 65536).all())
        assert torch.equal(out.view(n // len(blocks), -1), src.view(1, -1).expand(n // len(blocks), -1))
        ex.close()


def _engine_with(env: dict):
    import os
    from snappier_b200.batch import Engine
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return Engine(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_baseline_kernels_agree_with_fast_kernels(oracle, fixtures):
    """A/B: the simple v1 kernels (SNP_*_KERNEL=1) and the default kernels give identical bytes
    and statuses on corpus blocks, synthetic blocks and corrupted blocks."""
    from snappier_b200.batch import compress_many, decompress_many
    e1 = _engine_with({"SNP_DECOMP_KERNEL": "1", "SNP_COMP_KERNEL": "1"})
    e2 = _engine_with({})                            # default: tag-group engine, 4 KiB output window
    e4 = _engine_with({"SNP_V7_WINDOW": "2048"})     # the same engine with a 2 KiB window (40 warps per SM)
    e5 = _engine_with({"SNP_DECOMP_KERNEL": "8"})    # the challenger: lane-per-block engine
    _, blocks = _corpus_blocks(fixtures)
    blocks = blocks + H.synthetic_blocks(5150, 48)
    c1, s1 = compress_many(e1, blocks, 0)
    c2, s2 = compress_many(e2, blocks, 0)
    assert c1 == c2 and not s1.any() and not s2.any()
    for variant in ("6",):  # the challenger: plain 16-bit table entries in L2
        ev = _engine_with({"SNP_COMP_KERNEL": variant})
        for mode in (0, 1):
            cv, sv = compress_many(ev, blocks, mode)
            assert not sv.any() and cv == (c1 if mode == 0 else [oracle.compress(b, 1)[1] for b in blocks]), variant
        ev.close()
    rng = np.random.default_rng(17)
    items = list(c1)
    for c in c1[:40]:
        b = bytearray(c)
        b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        items.append(bytes(b))
    caps = [min(oracle.uncompressed_length(b)[1], 1 << 17) for b in items]
    d1, s1 = decompress_many(e1, items, caps)
    d2, s2 = decompress_many(e2, items, caps)
    assert np.array_equal(s1, s2) and d1 == d2
    assert d2[:len(blocks)] == blocks
    d4, s4 = decompress_many(e4, items, caps)
    assert np.array_equal(s4, s2) and d4 == d2
    for ex in (e5,):
        dx, sx = decompress_many(ex, items, caps)
        assert np.array_equal(sx, s1) and dx == d1
    # ragged / tiny / unaligned inputs through the ring's head-tail byte path
    small = [oracle.compress(b[:n])[1] for b in blocks[:8] for n in (0, 1, 5, 15, 16, 17, 31, 33, 255, 257, 511, 513, 1023, 1500)]
    d4s, s4s = decompress_many(e4, small)
    d2s, s2s = decompress_many(e2, small)
    assert d4s == d2s and not s4s.any()
    assert d2s == [oracle.decompress(c)[1] for c in small] and not s2s.any()
    # long literals at every source/destination alignment (the vectorised literal paths)
    rng2 = np.random.default_rng(23)
    lits = []
    for n in (127, 128, 129, 143, 144, 145, 300, 1000, 4097, 70001):
        for pad in (0, 1, 5, 15):
            raw = rng2.integers(0, 256, size=n + pad, dtype=np.uint8).tobytes()
            lits.append(oracle.compress(raw)[1])  # incompressible -> (pad-shifted) long literals
    for ex in (e1, e2, e4, e5):
        dl, sl = decompress_many(ex, lits)
        assert not sl.any() and dl == [oracle.decompress(c)[1] for c in lits]
    for ex in (e1, e2, e4, e5):
        ex.close()


@pytest.mark.parametrize("env", [{}, {"SNP_V7_WINDOW": "2048"}])
def test_ragged_blocks_work_stealing_and_dense_tags(oracle, fixtures, env):
    """The default kernel on thousands of ragged blocks that share 16-byte output vectors with their neighbours (the
    window flush must not touch a neighbour's bytes) and on a block of > 24 576 tiny tags."""
    from snappier_b200.batch import decompress_many
    e6 = _engine_with(env)
    _, blocks = _corpus_blocks(fixtures)
    blocks = blocks + H.synthetic_blocks(808, 60)
    rng = np.random.default_rng(5)
    raw = []
    for i in range(2500):
        b = blocks[i % len(blocks)]
        lo = int(rng.integers(0, 30000))
        raw.append(b[lo: lo + int(rng.integers(0, 35000))])
    # > 24 576 tags in one block: alternating 1-byte literals and 4-byte copies (budget = 768 groups of 32)
    dense = bytearray(oracle.varint_write(100000))
    dense += bytes([3 << 2]) + b"abcd"
    total = 4
    while total + 5 <= 100000:
        dense += bytes([0]) + b"x" + bytes([((4 - 1) << 2) | 2]) + (4).to_bytes(2, "little")
        total += 5
    dense += bytes([(100000 - total - 1) << 2]) + b"y" * (100000 - total)  # 1 byte left
    items = [oracle.compress(r)[1] for r in raw] + [bytes(dense)]
    got, st = decompress_many(e6, items)
    assert not st.any()
    assert got[:-1] == raw
    assert got[-1] == oracle.decompress(bytes(dense))[1] and len(got[-1]) == 100000
    e6.close()


@pytest.mark.parametrize("env", [{}, {"SNP_V7_WINDOW": "2048"}])
def test_handmade_tag_forms_and_fuzz(oracle, env):
    """The default kernel on the hand-assembled tag forms the emulator tests use (COPY4, literal-length forms, literals
    > 64 bytes at every position of a tag group, every offset 1..40 x lengths around 16 / 32 / 64) and on mutated
    blocks: status and bytes equal the oracle's."""
    from snappier_b200.batch import decompress_many
    from tests.helpers import handmade_tag_forms
    e6 = _engine_with(env)
    items = handmade_tag_forms()
    rng = np.random.default_rng(4)
    for b in list(items[5:40]):
        m = bytearray(b)
        for _ in range(3):
            m[int(rng.integers(0, len(m)))] = int(rng.integers(0, 256))
        items.append(bytes(m))
    caps = []
    for b in items:
        st, n = oracle.uncompressed_length(b)
        caps.append(min(n, 1 << 20) if st == 0 else 0)
    got, status = decompress_many(e6, items, caps)
    for i, b in enumerate(items):
        st, dec = oracle.decompress(b, cap=caps[i])
        assert status[i] == st, (i, status[i], st)
        assert got[i] == dec, i
    e6.close()


def test_host_mode_compress_many_chunks_in_flight(engine, oracle):
    """Host-mode batches flow through a 4-slot pipeline on 4 streams (16 384 items per chunk).  The L2-table compress
    kernels share one table buffer, so consecutive chunks must not overlap on the GPU: 50 000 small blocks (short
    kernels, many chunks) must still be bit-exact."""
    from snappier_b200.batch import compress_many, decompress_many
    blocks = H.synthetic_blocks(321, 50, size=3000)
    items = [blocks[i % 50][(i * 7) % 500:] for i in range(50000)]
    comp, st = compress_many(engine, items, 0)
    assert not st.any()
    want = {}
    for i in range(0, 50000, 97):
        key = (i % 50, (i * 7) % 500)
        if key not in want:
            want[key] = oracle.compress(items[i])[1]
        assert comp[i] == want[key], i
    back, st2 = decompress_many(engine, comp)
    assert not st2.any() and back == items


def test_single_call_api_is_thread_safe(oracle):
    """Snappy.* are static and re-entrant (SURVEY 8(b) "Threading"): concurrent calls from several
    threads, each on its lazily created thread-local context, all give oracle bytes."""
    import threading
    from snappier_b200 import snappy as S
    blocks = H.synthetic_blocks(4242, 24, size=20000)
    want = [oracle.compress(b)[1] for b in blocks]
    errors = []

    def worker(tid):
        try:
            for rep in range(3):
                for i in range(tid, len(blocks), 4):
                    c = S.compress_to_array(blocks[i])
                    assert c == want[i]
                    assert S.decompress_to_array(c) == blocks[i]
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors


def test_argument_errors(engine):
    """Call-level failures are negative codes, never crashes."""
    import ctypes as C
    from snappier_b200 import _native as N
    L = N.lib()
    w = C.c_size_t(0)
    buf = np.zeros(64, np.uint8)
    assert L.snp_compress(buf.ctypes.data, 8, buf.ctypes.data, 64, C.byref(w), 7) == N.E_INVALID_ARG  # bad hash mode
    assert L.snp_compress(buf.ctypes.data, 8, buf.ctypes.data + 4, 60, C.byref(w), 0) == N.E_OVERLAP
    assert L.snp_compress(None, 8, buf.ctypes.data, 64, C.byref(w), 0) == N.E_INVALID_ARG
    assert L.snp_decompress_batch(None, buf.ctypes.data, None, None, None, None, None, None, None, 3, 0, None) == N.E_INVALID_ARG
    assert L.snp_decompress_batch(None, None, None, None, None, None, None, None, None, 0, 0, None) == N.OK  # empty batch
    # an oversized item in a compress batch is reported per item, neighbours unaffected
    from snappier_b200.batch import pack
    items = [b"a" * 100, b"b" * 70000, b"c" * 100]
    base, off, ln = pack(items)
    out = np.zeros(3 * 80000, np.uint8)
    out_len, status = engine.compress_batch_host(base, off, ln, out, np.arange(3, dtype=np.uint64) * 80000,
                                                 np.full(3, 80000, np.uint32))
    assert status[0] == 0 and status[2] == 0 and status[1] == N.E_INVALID_ARG and out_len[1] == 0


def test_pack_batch_matches_concatenation(engine):
    """snp_pack_batch: slots with slack -> the items' bytes back to back (what a caller of Snappy.CompressToMemory gets
    per block), offsets = exclusive prefix sum, for ragged lengths (zero-length items, lengths around the 16-byte vector
    width, misaligned sources) and for more items than one scan CTA covers."""
    import torch
    rng = np.random.default_rng(21)
    dev = torch.device("cuda", 0)
    for n, maxlen in ((1, 100), (7, 40), (5000, 700), (70000, 90)):
        lens = rng.integers(0, maxlen, size=n).astype(np.int32)
        lens[rng.integers(0, n, size=max(1, n // 10))] = 0
        slack = rng.integers(0, 37, size=n).astype(np.int64)
        src_off = np.cumsum(lens.astype(np.int64) + slack) - lens  # item i sits behind its slack
        total_src = int(src_off[-1] + lens[-1]) + 64
        src = rng.integers(0, 256, size=total_src, dtype=np.uint8)
        want = b"".join(src[o:o + l].tobytes() for o, l in zip(src_off, lens))
        t_src = torch.from_numpy(src).to(dev)
        t_off = torch.from_numpy(src_off).to(dev)
        t_len = torch.from_numpy(lens).to(dev)
        off_q, tot_q = engine.pack_batch_device(t_src, t_off, t_len, None)
        assert int(tot_q) == len(want)
        dst = torch.full((len(want) + 32,), 0xEE, dtype=torch.uint8, device=dev)
        off, tot = engine.pack_batch_device(t_src, t_off, t_len, dst[16:])
        torch.cuda.synchronize()
        assert int(tot) == len(want)
        assert np.array_equal(off.cpu().numpy(), np.cumsum(lens.astype(np.int64)) - lens)
        got = dst.cpu().numpy()
        assert got[16:16 + len(want)].tobytes() == want
        assert (got[:16] == 0xEE).all() and (got[16 + len(want):] == 0xEE).all()


def test_find_match_length_kats_on_the_gpu(engine, kats):
    """The 45 known-answer vectors of SnappyCompressorTests.cs:10-81, run through the compress kernels' own device
    function (snp_find_match_length_batch) -- not only through whole-fragment parity."""
    import torch
    dev = torch.device("cuda", 0)
    buf, s1, s2, lim, want = bytearray(), [], [], [], []
    for k in kats["find_match_length"]:
        a, b, length = k["s1"].encode(), k["s2"].encode(), k["length"]
        while len(buf) % 4 != len(want) % 4:  # every alignment of the pair
            buf += b"\xee"
        s1.append(len(buf))
        buf += a
        s2.append(len(buf))
        buf += b + b"\0" * max(0, length - len(b))
        lim.append(s2[-1] + length)
        want.append(k["expected"])
        buf += b"\xdd" * 7
    assert len(want) == 45
    t = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    got = engine.find_match_length_batch_device(torch.tensor(list(buf) + [0] * 64, dtype=torch.uint8, device=dev), t(s1), t(s2), t(lim))
    torch.cuda.synchronize()
    assert got.cpu().tolist() == want


def _random_data_inputs(seed: int, count: int, big: int):
    """The input distribution of SnappyTests.cs:401-446 (RandomData): the first `big` inputs are 64..128 KiB of runs of
    arbitrary bytes, the rest 0..4 KiB of runs over a skewed small alphabet; run lengths skewed short."""
    rng = np.random.default_rng(seed)
    res = []
    for i in range(count):
        length = int(rng.integers(0, 4095)) if i >= big else 65536 + int(rng.integers(0, 65535))
        out = bytearray()
        while len(out) < length:
            run = 1
            if rng.integers(0, 9) == 0:
                run = int(rng.integers(0, max(1, (1 << int(rng.integers(0, 8))) - 1)))
            c = int(rng.integers(0, 255))
            if i >= big:
                c = int(rng.integers(0, max(1, (1 << int(rng.integers(0, 3))) - 1)))
            out += bytes([c]) * run
        res.append(bytes(out[:length]))
    return res


@pytest.mark.parametrize("kernel", ["7", "8"])
def test_random_data_distribution_round_trips(oracle, kernel):
    """SnappyTests.cs:401-446 on the GPU: 2 100 inputs of the RandomData distribution (100 above 64 KiB, through the
    single-call API with its fragment loop; 2 000 small ones as one batch), compressed bit-exactly like the oracle and
    decompressed back by the default and by the challenger decompress kernel."""
    from snappier_b200 import snappy as S
    from snappier_b200.batch import compress_many, decompress_many
    eng = _engine_with({"SNP_DECOMP_KERNEL": kernel})
    inputs = _random_data_inputs(301, 2100, 100)
    small = inputs[100:]
    comp, st = compress_many(eng, small, 0)
    assert not st.any()
    for i in range(0, len(small), 7):
        assert comp[i] == oracle.compress(small[i])[1], i
    back, st = decompress_many(eng, comp)
    assert not st.any() and back == small
    for d in inputs[:100:9]:  # multi-fragment inputs (Snappy.CompressToMemory / DecompressToMemory)
        c = S.compress_to_array(d)
        assert c == oracle.compress(d)[1]
        assert S.decompress_to_array(c) == d
    big_comp = [oracle.compress(d)[1] for d in inputs[:100]]
    back, st = decompress_many(eng, big_comp)  # > 64 KiB under one header, batched
    assert not st.any() and back == inputs[:100]
    eng.close()


@pytest.mark.parametrize("kernel", ["7", "8"])
def test_decode_at_scale_streams_the_engine_did_not_produce(oracle, kernel):
    """2^14 blocks of bench.py's config-2 mixture compressed by the ORACLE (both hash modes) and by Google's encoder
    (pyarrow) -- not by the engine's own compressor -- decode bit-exactly in one batched call."""
    import torch
    import bench as B
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")
    n = 1 << 14
    corpus = {k: torch.from_numpy(v) for k, v in B.load_corpus().items()}
    raw = torch.cat([B.make_blocks(torch, corpus, b0, 4096, torch.device("cpu")) for b0 in range(0, n, 4096)]).numpy()
    r_off = np.arange(n, dtype=np.uint64) * B.BLOCK
    r_len = np.full(n, B.BLOCK, np.uint32)
    s_off = np.arange(n, dtype=np.uint64) * B.PITCH
    s_cap = np.full(n, B.PITCH, np.uint32)
    eng = _engine_with({"SNP_DECOMP_KERNEL": kernel})
    for producer in ("oracle-crc32c", "oracle-mul", "google"):
        slots = np.zeros(n * B.PITCH, np.uint8)
        if producer == "google":
            lens = np.zeros(n, np.uint32)
            for i in range(0, n, 4):  # every 4th block through pyarrow (one thread), the rest left empty
                c = codec.compress(raw[i].tobytes()).to_pybytes()
                slots[i * B.PITCH: i * B.PITCH + len(c)] = np.frombuffer(c, np.uint8)
                lens[i] = len(c)
            sel = np.arange(0, n, 4)
        else:
            bad, lens, _ = oracle.compress_batch(raw.reshape(-1), r_off, r_len, slots, s_off, s_cap,
                                                 0 if producer == "oracle-crc32c" else 1, 8)
            assert bad == 0
            sel = np.arange(n)
        out = np.zeros(len(sel) * B.BLOCK, np.uint8)
        ol, st = eng.decompress_batch_host(slots, s_off[sel], lens[sel], out, np.arange(len(sel), dtype=np.uint64) * B.BLOCK,
                                           np.full(len(sel), B.BLOCK, np.uint32))
        assert not st.any() and (ol == B.BLOCK).all(), producer
        assert np.array_equal(out.reshape(len(sel), B.BLOCK), raw[sel]), producer
    eng.close()


def test_single_call_api_from_many_threads(oracle):
    """The single-call API behind Snappy.Compress / Decompress is called from arbitrary thread-pool threads: all of them
    share ONE default context per device (calls serialise on its mutex), results stay bit-exact and independent."""
    import threading
    from snappier_b200 import snappy as S
    blocks = H.synthetic_blocks(909, 24, size=20000) + [b"", b"x", b"thread " * 5000]
    want = [oracle.compress(b)[1] for b in blocks]
    errors = []

    def worker(tid):
        try:
            for rep in range(3):
                for i in range(tid % 3, len(blocks), 3):
                    c = S.compress_to_array(blocks[i])
                    assert c == want[i], (tid, i)
                    assert S.decompress_to_array(c) == blocks[i], (tid, i)
                    assert S.get_uncompressed_length(c) == len(blocks[i])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(12)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
