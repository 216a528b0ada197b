"""Python mirror of Snappier's static `Snappy` block facade, bound to the C ABI.

Same method names (snake_case), argument meaning and error behaviour as
/root/reference/Snappier/Snappy.cs -- each method cites the reference member it
mirrors -- so the parity tests read like the reference's own.  All work happens in
libsnappier_b200.so (CUDA, sm_100a); there is no Python or CPU implementation.

.NET exception -> Python exception:
    ArgumentException          -> ArgumentException(ValueError)
    InvalidDataException       -> InvalidDataException(ValueError)
    InvalidOperationException  -> InvalidOperationException(RuntimeError)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


class ArgumentException(ValueError):
    pass


class InvalidDataException(ValueError):
    pass


class InvalidOperationException(RuntimeError):
    pass


HASH_CRC32C = N.HASH_CRC32C
HASH_MUL = N.HASH_MUL

#: hash variant used by compress(); Snappier on x64/.NET 8+ uses CRC32C (HashTable.cs:109-117)
default_hash_mode = HASH_CRC32C


def _ro(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf, dtype=np.uint8)
    return np.frombuffer(buf, dtype=np.uint8)


def _ptr(a: np.ndarray):
    return a.ctypes.data if a.size else None


def _raise_for_status(st: int, where: str) -> None:
    """Status -> exception, as SURVEY.md section 8(b) / include/snappier_b200.h list them."""
    if st == N.OK:
        return
    if st == N.OUTPUT_TOO_SMALL:
        raise ArgumentException("Output buffer is too small.")  # ThrowHelper.cs:18-19
    if st == N.INVALID_LENGTH:
        if where == "decompress":
            raise InvalidOperationException("Invalid stream length")  # SnappyDecompressor.cs:53-56
        raise InvalidDataException("Invalid stream length")  # VarIntEncoding.Read.cs:20
    if st == N.INCOMPLETE:
        raise InvalidDataException("Incomplete Snappy block.")  # ThrowHelper.cs:27-28
    if st == N.INVALID_COPY_OFFSET:
        raise InvalidDataException("Invalid copy offset")  # SnappyDecompressor.cs:600
    if st == N.DATA_TOO_LONG:
        raise InvalidDataException("Data too long")  # SnappyDecompressor.cs:572,605
    if st == N.E_OVERLAP:
        raise InvalidOperationException("Input and output spans must not overlap.")  # SnappyCompressor.cs:29
    N.check_call(st, where)
    raise RuntimeError(f"{where}: unexpected status {st}")


def get_max_compressed_length(input_length: int) -> int:
    """Snappy.GetMaxCompressedLength (Snappy.cs:20-24)."""
    return N.lib().snp_get_max_compressed_length(input_length)


def try_compress(input, output: np.ndarray, hash_mode: int | None = None) -> tuple[bool, int]:
    """Snappy.TryCompress (Snappy.cs:55-67) -> (success, bytesWritten).

    `output` is a writable uint8 numpy array (the Span<byte>)."""
    a = _ro(input)
    w = C.c_size_t(0)
    st = N.lib().snp_compress(_ptr(a), a.size, _ptr(output), output.size, C.byref(w),
                              default_hash_mode if hash_mode is None else hash_mode)
    if st == N.OUTPUT_TOO_SMALL:
        return False, 0
    _raise_for_status(st, "compress")
    return True, w.value


def compress(input, output: np.ndarray, hash_mode: int | None = None) -> int:
    """Snappy.Compress(ReadOnlySpan<byte>, Span<byte>) (Snappy.cs:37-45)."""
    ok, n = try_compress(input, output, hash_mode)
    if not ok:
        raise ArgumentException("Output buffer is too small.")
    return n


def compress_to_memory(input, hash_mode: int | None = None) -> np.ndarray:
    """Snappy.CompressToMemory (Snappy.cs:99-113): returns the owned buffer, trimmed."""
    a = _ro(input)
    buf = np.empty(get_max_compressed_length(a.size), np.uint8)
    ok, n = try_compress(a, buf, hash_mode)
    if not ok:
        raise InvalidOperationException()  # Snappy.cs:109 "should be unreachable"
    return buf[:n]


def compress_to_array(input, hash_mode: int | None = None) -> bytes:
    """Snappy.CompressToArray (Snappy.cs:123-132)."""
    return compress_to_memory(input, hash_mode).tobytes()


def _segments(segments):
    """list of byte-likes -> (kept-alive arrays, void*[n], size_t[n])"""
    arrs = [_ro(x) for x in segments]
    ptrs = (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data if a.size else None for a in arrs])
    lens = (C.c_size_t * max(len(arrs), 1))(*[a.size for a in arrs])
    return arrs, ptrs, lens


def compress_sequence(segments, hash_mode: int | None = None) -> bytes:
    """Snappy.Compress(ReadOnlySequence<byte>, IBufferWriter<byte>) (Snappy.cs:82-89) -> the bytes handed to the
    writer.  `segments` is the sequence's list of memory segments: the reference cuts its fragments along them
    (SnappyCompressor.cs:103-143), so the output depends on the segmentation, not only on the bytes."""
    arrs, ptrs, lens = _segments(segments)
    total = sum(a.size for a in arrs)
    if total > 0xFFFFFFFF:
        raise ArgumentException("input is larger than the maximum size of 4294967295 bytes.")
    # every fragment may be shorter than 64 KiB here, so size for the worst case per segment boundary
    cap = get_max_compressed_length(total) + 64 * (len(arrs) + total // 32768 + 1)
    out = np.empty(cap, np.uint8)
    w = C.c_size_t(0)
    st = N.lib().snp_compress_sequence(ptrs, lens, len(arrs), _ptr(out), out.size, C.byref(w),
                                       default_hash_mode if hash_mode is None else hash_mode)
    _raise_for_status(st, "compress")
    return out[: w.value].tobytes()


def decompress_sequence(segments) -> bytes:
    """Snappy.DecompressToMemory(ReadOnlySequence<byte>) (Snappy.cs:246-261): one block split into segments."""
    arrs, ptrs, lens = _segments(segments)
    head = _ro(b"".join(a[:5].tobytes() for a in arrs)[:5])  # the varint prefix may straddle segments
    v = C.c_uint32(0)
    n = v.value if N.lib().snp_uncompressed_length(_ptr(head), head.size, C.byref(v)) == N.OK else 0
    n = v.value
    out = np.empty(max(n, 1), np.uint8)
    w = C.c_size_t(0)
    st = N.lib().snp_decompress_sequence(ptrs, lens, len(arrs), _ptr(out), n, C.byref(w))
    _raise_for_status(st, "decompress")
    return out[: w.value].tobytes()


def get_uncompressed_length(input) -> int:
    """Snappy.GetUncompressedLength (Snappy.cs:142-143)."""
    a = _ro(input)
    v = C.c_uint32(0)
    st = N.lib().snp_uncompressed_length(_ptr(a), a.size, C.byref(v))
    _raise_for_status(st, "get_uncompressed_length")
    return v.value


def try_decompress(input, output: np.ndarray) -> tuple[bool, int]:
    """Snappy.TryDecompress (Snappy.cs:172-186) -> (success, bytesWritten).

    Data errors raise; a too-small output returns (False, bytes copied)."""
    a = _ro(input)
    w = C.c_size_t(0)
    st = N.lib().snp_decompress(_ptr(a), a.size, _ptr(output), output.size, C.byref(w))
    if st == N.OUTPUT_TOO_SMALL:
        return False, w.value
    _raise_for_status(st, "decompress")
    return True, w.value


def decompress(input, output: np.ndarray) -> int:
    """Snappy.Decompress(ReadOnlySpan<byte>, Span<byte>) (Snappy.cs:153-162)."""
    ok, n = try_decompress(input, output)
    if not ok:
        raise ArgumentException("Output buffer is too small.")
    return n


def decompress_to_memory(input) -> np.ndarray:
    """Snappy.DecompressToMemory(ReadOnlySpan<byte>) (Snappy.cs:223-235)."""
    a = _ro(input)
    v = C.c_uint32(0)
    st = N.lib().snp_uncompressed_length(_ptr(a), a.size, C.byref(v))
    # DecompressToMemory goes through SnappyDecompressor.Decompress: a truncated
    # prefix is "Incomplete Snappy block.", an overflowing one InvalidOperationException.
    out = np.empty(v.value if st == N.OK else 0, np.uint8)
    ok, n = try_decompress(a, out)
    assert ok
    return out[:n]


def decompress_to_array(input) -> bytes:
    """Snappy.DecompressToArray (Snappy.cs:273-282)."""
    length = get_uncompressed_length(input)
    out = np.empty(length, np.uint8)
    decompress(input, out)
    return out.tobytes()
