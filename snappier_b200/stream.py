"""One-shot Snappy *framing format* calls over the C ABI.

`frame_compress(data)` returns exactly the bytes `new SnappyStream(s, CompressionMode.Compress)`
produces for one Write of `data` followed by Dispose (SnappyStreamCompressor.cs:15-18,166-261);
`frame_decompress(stream)` is reading a `SnappyStream(s, CompressionMode.Decompress)` to the end
(SnappyStreamDecompressor.cs:38-208).  The .NET `Stream` plumbing itself (SnappyStream.cs) is out
of scope and stays C#.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .snappy import (ArgumentException, InvalidDataException, _ptr, _raise_for_status, _ro)  # noqa: F401


def _raise(st: int, where: str) -> None:
    if st == N.UNKNOWN_CHUNK_TYPE:
        raise InvalidDataException("Unknown chunk type")  # SnappyStreamDecompressor.cs:182-185
    if st == N.CRC_MISMATCH:
        raise InvalidDataException("Chunk CRC mismatch.")  # SnappyStreamDecompressor.cs:127-131
    _raise_for_status(st, where)


def frame_max_compressed_length(n: int) -> int:
    return int(N.lib().snp_frame_max_compressed_length(n))


def frame_compress(data, hash_mode: int = N.HASH_CRC32C) -> bytes:
    a = _ro(data)
    out = np.empty(frame_max_compressed_length(a.size), np.uint8)
    w = C.c_size_t(0)
    st = N.lib().snp_frame_compress(_ptr(a), a.size, _ptr(out), out.size, C.byref(w), hash_mode)
    _raise(st, "frame_compress")
    return out[: w.value].tobytes()


def frame_uncompressed_length(stream) -> int:
    a = _ro(stream)
    v = C.c_uint64(0)
    st = N.lib().snp_frame_uncompressed_length(_ptr(a), a.size, C.byref(v))
    _raise(st, "frame_uncompressed_length")
    return v.value


def frame_decompress(stream) -> bytes:
    a = _ro(stream)
    out = np.empty(max(frame_uncompressed_length(a), 1), np.uint8)
    w = C.c_size_t(0)
    st = N.lib().snp_frame_decompress(_ptr(a), a.size, _ptr(out), out.size, C.byref(w))
    _raise(st, "frame_decompress")
    return out[: w.value].tobytes()


def crc32c_batch(engine, base: np.ndarray, off, length, masked: bool = True) -> np.ndarray:
    """Crc32CAlgorithm.Compute (+ ApplyMask) per item on host arrays."""
    off = np.ascontiguousarray(off, np.uint64)
    length = np.ascontiguousarray(length, np.uint32)
    crc = np.zeros(off.size, np.uint32)
    rc = N.lib().snp_crc32c_batch(engine._ctx, _ptr(base), _ptr(off), _ptr(length), _ptr(crc), off.size,
                                  1 if masked else 0, N.MEM_HOST, None)
    N.check_call(rc, "snp_crc32c_batch")
    return crc
