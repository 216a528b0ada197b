"""GPU parity for the framing-format row (BASELINE config 4): CRC32C kernel,
frame_compress / frame_decompress through the C ABI against the reference's framed goldens and
the oracle-side checker."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_crc32c_kats_and_random_items(engine, oracle, kats):
    """Crc32CAlgorithmTests.cs:7-24 + ragged, unaligned items up to > 64 KiB."""
    from snappier_b200.batch import pack
    from snappier_b200.stream import crc32c_batch
    rng = np.random.default_rng(4)
    items = [k["ascii"].encode() for k in kats["crc32c"]]
    for n in (0, 1, 2, 3, 4, 5, 7, 63, 64, 65, 127, 128, 129, 131, 255, 256, 1000, 4095, 4096, 4097, 65535, 65536, 65537, 200001):
        items.append(rng.integers(0, 256, size=n, dtype=np.uint8).tobytes())
    base, off, ln = pack(items)  # consecutive packing -> every alignment class occurs
    got = crc32c_batch(engine, base, off, ln, masked=False)
    assert [int(x) for x in got[:4]] == [k["crc"] for k in kats["crc32c"]]
    assert [int(x) for x in got] == [oracle.crc32c(b) for b in items]
    got_m = crc32c_batch(engine, base, off, ln, masked=True)
    assert [int(x) for x in got_m] == [oracle.crc32c_masked(b) for b in items]


def test_golden_framed_streams_byte_exact(oracle, fixtures):
    """MUL-hash frame_compress reproduces the reference's framed goldens byte for byte; both
    goldens decode to the corpus with every chunk CRC verified on the GPU."""
    from snappier_b200 import stream as S
    html4 = fixtures["corpus/html_x_4"]
    alice_crlf = fixtures["corpus/alice29.txt"].replace(b"\n", b"\r\n")
    for raw, name in ((html4, "html_x_4"), (alice_crlf, "alice29")):
        gold = fixtures[f"framed/{name}.snappy"]
        assert S.frame_compress(raw, 1) == gold
        assert S.frame_decompress(gold) == raw
        assert S.frame_uncompressed_length(gold) == len(raw)


def test_frame_round_trip_corpus_and_edge_sizes(oracle, fixtures):
    from snappier_b200 import stream as S
    rng = np.random.default_rng(8)
    cases = [fixtures["corpus/" + f] for f in ("fireworks.jpeg", "urls.10K", "kppkn.gtb", "paper-100k.pdf")]
    cases += [b"", b"a", b"ab" * 40000, rng.integers(0, 256, size=256, dtype=np.uint8).tobytes(),
              rng.integers(0, 256, size=65536 * 3 + 17, dtype=np.uint8).tobytes(), b"x" * 65536, b"y" * 65537]
    for d in cases:
        for mode in (0, 1):
            f = S.frame_compress(d, mode)
            assert f == oracle.frame_compress(d, mode)
            assert S.frame_decompress(f) == d
    # uncompressed-chunk size == 10 + 8 + 256 (SnappyStreamCompressorTests.cs:7-46)
    assert len(S.frame_compress(cases[7])) == 10 + 8 + 256


def test_frame_reader_semantics(oracle, fixtures):
    """SnappyStreamDecompressor.cs:180-199: skippable chunks and stream-identifier content are
    ignored, reserved unskippable types and CRC mismatches throw; truncated streams are incomplete."""
    from snappier_b200 import stream as S
    d = fixtures["corpus/html"]
    f = bytearray(oracle.frame_compress(d))
    skip = bytes([0x80, 3, 0, 0, 1, 2, 3]) + bytes([0xfe, 2, 0, 0, 0, 0])  # skippable + padding chunks
    f2 = bytes(f[:10]) + skip + bytes(f[10:]) + skip
    assert S.frame_decompress(f2) == d
    f3 = bytearray(f)
    f3[4:10] = b"NOTSNP"  # stream identifier content is not validated
    assert S.frame_decompress(bytes(f3)) == d
    f4 = bytes(f[:10]) + bytes([0x02, 1, 0, 0, 9]) + bytes(f[10:])
    with pytest.raises(S.InvalidDataException, match="Unknown chunk type"):
        S.frame_decompress(f4)
    f5 = bytearray(f)
    f5[10 + 4] ^= 0x01  # first chunk's CRC
    with pytest.raises(S.InvalidDataException, match="CRC"):
        S.frame_decompress(bytes(f5))
    assert oracle.frame_decompress(bytes(f5))[0] == oracle.CRC_MISMATCH
    f6 = bytearray(f)
    f6[10 + 8 + 40] ^= 0x55  # payload corruption: block error or CRC mismatch, same as the checker
    st, _ = oracle.frame_decompress(bytes(f6))
    assert st != 0
    with pytest.raises(S.InvalidDataException):
        S.frame_decompress(bytes(f6))
    with pytest.raises(S.InvalidDataException, match="Incomplete"):
        S.frame_decompress(bytes(f[:-5]))
