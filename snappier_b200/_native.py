"""ctypes loader for libsnappier_b200.so (the C ABI in include/snappier_b200.h).

There is no fallback: if the library is missing or cannot be loaded, importing
the compute API raises.  Only the C-ABI symbols are bound here -- no torch types
cross this boundary (pointers are passed as integers).
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_PKG, "libsnappier_b200.so")

# enum snp_status / snp_error / snp_hash_mode / snp_mem_kind
OK, OUTPUT_TOO_SMALL, INVALID_LENGTH, INCOMPLETE, INVALID_COPY_OFFSET, DATA_TOO_LONG = range(6)
UNKNOWN_CHUNK_TYPE, CRC_MISMATCH = 6, 7  # framing format only
E_CUDA, E_INVALID_ARG, E_NO_DEVICE, E_OVERLAP = -1, -2, -3, -4
HASH_CRC32C, HASH_MUL = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
BLOCK_SIZE = 65536

# every symbol include/snappier_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "snp_abi_version", "snp_status_string", "snp_last_error",
    "snp_max_compressed_length", "snp_get_max_compressed_length", "snp_uncompressed_length",
    "snp_create", "snp_destroy", "snp_ctx_device", "snp_ctx_launch_count",
    "snp_compress", "snp_decompress", "snp_compress_sequence", "snp_decompress_sequence",
    "snp_compress_batch", "snp_decompress_batch", "snp_uncompressed_length_batch",
    "snp_frame_max_compressed_length", "snp_frame_compress", "snp_frame_uncompressed_length",
    "snp_frame_decompress", "snp_crc32c_batch", "snp_pack_batch", "snp_find_match_length_batch", "snp_diag_random_reads",
]

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise NativeLibraryError(
            f"{SO_PATH} not found: build it with `python -m snappier_b200.build` "
            "(there is no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32
    L.snp_abi_version.restype = C.c_int
    L.snp_status_string.argtypes = [C.c_int]
    L.snp_status_string.restype = C.c_char_p
    L.snp_last_error.restype = C.c_char_p
    L.snp_max_compressed_length.argtypes = [i32]
    L.snp_max_compressed_length.restype = i32
    L.snp_get_max_compressed_length.argtypes = [i32]
    L.snp_get_max_compressed_length.restype = i32
    L.snp_uncompressed_length.argtypes = [vp, sz, C.POINTER(u32)]
    L.snp_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.snp_destroy.argtypes = [vp]
    L.snp_destroy.restype = None
    L.snp_ctx_device.argtypes = [vp]
    L.snp_ctx_launch_count.argtypes = [vp]
    L.snp_ctx_launch_count.restype = C.c_uint64
    L.snp_compress.argtypes = [vp, sz, vp, sz, C.POINTER(sz), u32]
    L.snp_decompress.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
    L.snp_compress_sequence.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), u32]
    L.snp_decompress_sequence.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
    L.snp_compress_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, u32, C.c_int, vp]
    L.snp_decompress_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, C.c_int, vp]
    L.snp_uncompressed_length_batch.argtypes = [vp, vp, vp, vp, vp, vp, sz, C.c_int, vp]
    L.snp_frame_max_compressed_length.argtypes = [sz]
    L.snp_frame_max_compressed_length.restype = sz
    L.snp_frame_compress.argtypes = [vp, sz, vp, sz, C.POINTER(sz), u32]
    L.snp_frame_uncompressed_length.argtypes = [vp, sz, C.POINTER(C.c_uint64)]
    L.snp_frame_decompress.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
    L.snp_crc32c_batch.argtypes = [vp, vp, vp, vp, vp, sz, C.c_int, C.c_int, vp]
    L.snp_pack_batch.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp, vp]
    L.snp_find_match_length_batch.argtypes = [vp, vp, vp, vp, vp, vp, sz, vp]
    L.snp_diag_random_reads.argtypes = [vp, vp, sz, u32, u32, vp, vp]
    _lib = L
    return L


def status_string(st: int) -> str:
    return lib().snp_status_string(st).decode()


def last_error() -> str:
    return lib().snp_last_error().decode()


def check_call(rc: int, what: str) -> None:
    """Raise for call-level failures (negative codes).  Per-item statuses are not errors here."""
    if rc < 0:
        detail = last_error() if rc == E_CUDA or rc == E_NO_DEVICE else ""
        raise NativeLibraryError(f"{what}: {status_string(rc)} ({rc}) {detail}".rstrip())
