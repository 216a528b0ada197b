"""ctypes binding of the CPU oracle (oracle/libsnappy_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg -- never from snappier_b200/.
See oracle/snappy_oracle.h for the parity status of each function.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libsnappy_oracle.so")

OK, OUTPUT_TOO_SMALL, INVALID_LENGTH, INCOMPLETE, INVALID_COPY_OFFSET, DATA_TOO_LONG = range(6)
HASH_CRC32C, HASH_MUL = 0, 1
BLOCK_SIZE = 65536


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, seconds)."""
    src = os.path.join(_DIR, "snappy_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_DIR, "snappy_oracle.h")))
    if force or stale:
        subprocess.check_call(["make", "-C", _DIR, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p, u32p, u64p, i32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32))
        L.orc_max_compressed_length.argtypes = [C.c_int32]
        L.orc_get_max_compressed_length.argtypes = [C.c_int32]
        L.orc_varint_write.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
        L.orc_varint_read.argtypes = [C.c_void_p, C.c_size_t, u32p, C.POINTER(C.c_int)]
        L.orc_table_size.argtypes = [C.c_int]
        L.orc_table_hash.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.orc_table_hash.restype = C.c_uint32
        L.orc_table_hash_fast.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.orc_table_hash_fast.restype = C.c_uint32
        L.orc_find_match_length.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_compress_fragment.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.orc_compress_fragment.restype = C.c_size_t
        L.orc_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_size_t), C.c_int]
        L.orc_uncompressed_length.argtypes = [C.c_void_p, C.c_size_t, u32p]
        L.orc_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_size_t)]
        L.orc_crc32c.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        L.orc_crc32c.restype = C.c_uint32
        L.orc_crc32c_mask.argtypes = [C.c_uint32]
        L.orc_crc32c_mask.restype = C.c_uint32
        batch = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                 C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_compress_batch.argtypes = batch + [C.c_int, C.c_int]
        L.orc_decompress_batch.argtypes = batch + [C.c_int]
        _lib = L
    return _lib


def _buf(b) -> np.ndarray:
    a = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    return np.ascontiguousarray(a, dtype=np.uint8)


def max_compressed_length(n: int) -> int:
    return lib().orc_max_compressed_length(n)


def get_max_compressed_length(n: int) -> int:
    return lib().orc_get_max_compressed_length(n)


def varint_write(v: int) -> bytes:
    out = np.zeros(8, np.uint8)
    n = lib().orc_varint_write(out.ctypes.data, 8, v)
    return out[:n].tobytes()


def varint_read(b) -> tuple[int, int, int]:
    """-> (status, value, bytes consumed)"""
    a = _buf(b)
    v, used = C.c_uint32(0), C.c_int(0)
    st = lib().orc_varint_read(a.ctypes.data if a.size else None, a.size, C.byref(v), C.byref(used))
    return st, v.value, used.value


def find_match_length(buf, s1: int, s2: int, s2_limit: int) -> int:
    a = _buf(buf)
    base = a.ctypes.data
    return lib().orc_find_match_length(base + s1, base + s2, base + s2_limit)


def compress(data, hash_mode: int = HASH_CRC32C, cap: int | None = None) -> tuple[int, bytes]:
    """-> (status, compressed bytes)"""
    a = _buf(data)
    if cap is None:
        cap = get_max_compressed_length(a.size)
    out = np.zeros(max(cap, 1), np.uint8)
    w = C.c_size_t(0)
    st = lib().orc_compress(a.ctypes.data if a.size else None, a.size, out.ctypes.data, cap,
                            C.byref(w), hash_mode)
    return st, out[: w.value].tobytes()


def uncompressed_length(data) -> tuple[int, int]:
    a = _buf(data)
    v = C.c_uint32(0)
    st = lib().orc_uncompressed_length(a.ctypes.data if a.size else None, a.size, C.byref(v))
    return st, v.value


def decompress(data, cap: int | None = None) -> tuple[int, bytes]:
    """-> (status, decompressed bytes)"""
    a = _buf(data)
    if cap is None:
        st, cap = uncompressed_length(a)
        if st != OK:
            cap = 0
    out = np.zeros(max(cap, 1), np.uint8)
    w = C.c_size_t(0)
    st = lib().orc_decompress(a.ctypes.data if a.size else None, a.size, out.ctypes.data, cap,
                              C.byref(w))
    return st, out[: w.value].tobytes()


def crc32c(data, crc: int = 0) -> int:
    a = _buf(data)
    return lib().orc_crc32c(crc, a.ctypes.data if a.size else None, a.size)


def crc32c_masked(data) -> int:
    return lib().orc_crc32c_mask(crc32c(data))


def compress_batch(in_base: np.ndarray, in_off, in_len, out_base: np.ndarray, out_off, out_cap,
                   hash_mode: int = HASH_CRC32C, threads: int = 1):
    """numpy-array batch driver; returns (first_bad_status, out_len u32[N], status i32[N])."""
    n = len(in_off)
    in_off = np.ascontiguousarray(in_off, np.uint64)
    in_len = np.ascontiguousarray(in_len, np.uint32)
    out_off = np.ascontiguousarray(out_off, np.uint64)
    out_cap = np.ascontiguousarray(out_cap, np.uint32)
    out_len = np.zeros(n, np.uint32)
    status = np.zeros(n, np.int32)
    bad = lib().orc_compress_batch(in_base.ctypes.data, in_off.ctypes.data, in_len.ctypes.data,
                                   out_base.ctypes.data, out_off.ctypes.data, out_cap.ctypes.data,
                                   out_len.ctypes.data, status.ctypes.data, n, hash_mode, threads)
    return bad, out_len, status


def decompress_batch(in_base: np.ndarray, in_off, in_len, out_base: np.ndarray, out_off, out_cap,
                     threads: int = 1):
    n = len(in_off)
    in_off = np.ascontiguousarray(in_off, np.uint64)
    in_len = np.ascontiguousarray(in_len, np.uint32)
    out_off = np.ascontiguousarray(out_off, np.uint64)
    out_cap = np.ascontiguousarray(out_cap, np.uint32)
    out_len = np.zeros(n, np.uint32)
    status = np.zeros(n, np.int32)
    bad = lib().orc_decompress_batch(in_base.ctypes.data, in_off.ctypes.data, in_len.ctypes.data,
                                     out_base.ctypes.data, out_off.ctypes.data, out_cap.ctypes.data,
                                     out_len.ctypes.data, status.ctypes.data, n, threads)
    return bad, out_len, status


# ---- framing format checker (Python over the C primitives; SnappyStreamCompressor.cs:194-261) ----
STREAM_ID = bytes([0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59])
UNKNOWN_CHUNK_TYPE, CRC_MISMATCH = 6, 7


def frame_compress(data: bytes, hash_mode: int = HASH_CRC32C) -> bytes:
    out = bytearray(STREAM_ID)
    for i in range(0, len(data), BLOCK_SIZE):
        chunk = data[i:i + BLOCK_SIZE]
        st, c = compress(chunk, hash_mode)
        assert st == OK
        ctype, payload = (0x00, c) if len(c) < len(chunk) else (0x01, chunk)
        out += bytes([ctype]) + (len(payload) + 4).to_bytes(3, "little") + crc32c_masked(chunk).to_bytes(4, "little") + payload
    return bytes(out)


def frame_decompress(stream: bytes) -> tuple[int, bytes]:
    out, i = bytearray(), 0
    while i < len(stream):
        if len(stream) - i < 4:
            return INCOMPLETE, b""
        t, n = stream[i], int.from_bytes(stream[i + 1:i + 4], "little")
        i += 4
        if len(stream) - i < n:
            return INCOMPLETE, b""
        body = stream[i:i + n]
        i += n
        if t in (0, 1):
            if n < 4:
                return INCOMPLETE, b""
            crc = int.from_bytes(body[:4], "little")
            if t == 0:
                st, raw = decompress(body[4:])
                if st != OK:
                    return st, b""
            else:
                raw = body[4:]
            if crc32c_masked(raw) != crc:
                return CRC_MISMATCH, b""
            out += raw
        elif t < 0x80:
            return UNKNOWN_CHUNK_TYPE, b""
    return OK, bytes(out)
