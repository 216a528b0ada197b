"""Batched block API over the C ABI (snp_*_batch in include/snappier_b200.h).

The reference has no batch call; item i of a batch is exactly one
`Snappy.Compress` / `Snappy.Decompress` of an independent block
(SnappyCompressor.cs:40-44: fragments are independent by construction).

`Engine` wraps one snp_ctx (one GPU).  Buffers are passed either as torch CUDA
tensors (device mode: enqueued on a stream, no copies) or as numpy arrays (host
mode: staged through the library, synchronous).  torch is only the allocator /
stream provider here; every byte of codec work happens in libsnappier_b200.so.

Offsets are 64-bit, lengths/capacities/status 32-bit; torch tensors use int64 /
int32 with the same bit patterns as the ABI's uint64 / uint32.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def _np_ptr(a: np.ndarray):
    return a.ctypes.data if a.size else None


class Engine:
    def __init__(self, device: int = 0):
        self._ctx = C.c_void_p(None)
        rc = N.lib().snp_create(device, C.byref(self._ctx))
        N.check_call(rc, "snp_create")
        self.device = device

    def close(self) -> None:
        if self._ctx:
            N.lib().snp_destroy(self._ctx)
            self._ctx = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(N.lib().snp_ctx_launch_count(self._ctx))

    # ------------------------------------------------------------- device mode
    @staticmethod
    def _check_dev(t, dtype_size: int, name: str):
        if not t.is_cuda or not t.is_contiguous() or t.element_size() != dtype_size:
            raise ValueError(f"{name}: need a contiguous CUDA tensor with {dtype_size}-byte elements")
        return t.data_ptr()

    def compress_batch_device(self, in_base, in_off, in_len, out_base, out_off, out_cap, out_len,
                              status, hash_mode: int = N.HASH_CRC32C, stream: int | None = None) -> None:
        """Enqueue one batched compress; tensors: u8, i64[N], i32[N], u8, i64[N], i32[N], i32[N], i32[N]."""
        n = in_off.numel()
        rc = N.lib().snp_compress_batch(
            self._ctx, self._check_dev(in_base, 1, "in_base"), self._check_dev(in_off, 8, "in_off"),
            self._check_dev(in_len, 4, "in_len"), self._check_dev(out_base, 1, "out_base"),
            self._check_dev(out_off, 8, "out_off"), self._check_dev(out_cap, 4, "out_cap"),
            self._check_dev(out_len, 4, "out_len"), self._check_dev(status, 4, "status"), n, hash_mode,
            N.MEM_DEVICE, stream)
        N.check_call(rc, "snp_compress_batch")

    def decompress_batch_device(self, in_base, in_off, in_len, out_base, out_off, out_cap, out_len,
                                status, stream: int | None = None) -> None:
        n = in_off.numel()
        rc = N.lib().snp_decompress_batch(
            self._ctx, self._check_dev(in_base, 1, "in_base"), self._check_dev(in_off, 8, "in_off"),
            self._check_dev(in_len, 4, "in_len"), self._check_dev(out_base, 1, "out_base"),
            self._check_dev(out_off, 8, "out_off"), self._check_dev(out_cap, 4, "out_cap"),
            self._check_dev(out_len, 4, "out_len"), self._check_dev(status, 4, "status"), n,
            N.MEM_DEVICE, stream)
        N.check_call(rc, "snp_decompress_batch")

    def uncompressed_length_batch_device(self, in_base, in_off, in_len, ulen, status,
                                         stream: int | None = None) -> None:
        n = in_off.numel()
        rc = N.lib().snp_uncompressed_length_batch(
            self._ctx, self._check_dev(in_base, 1, "in_base"), self._check_dev(in_off, 8, "in_off"),
            self._check_dev(in_len, 4, "in_len"), self._check_dev(ulen, 4, "ulen"),
            self._check_dev(status, 4, "status"), n, N.MEM_DEVICE, stream)
        N.check_call(rc, "snp_uncompressed_length_batch")

    def pack_batch_device(self, src_base, src_off, length, dst_base=None, stream: int | None = None):
        """Dense packing of a batch (slots with slack -> exactly the items' bytes, in order).  Returns (dst_off i64[N],
        total i64[1]) on the device; with dst_base=None only the offsets / total are computed (size query)."""
        import torch
        n = src_off.numel()
        dst_off = torch.empty(n, dtype=torch.int64, device=src_off.device)
        total = torch.zeros(1, dtype=torch.int64, device=src_off.device)
        rc = N.lib().snp_pack_batch(
            self._ctx, self._check_dev(src_base, 1, "src_base"), self._check_dev(src_off, 8, "src_off"),
            self._check_dev(length, 4, "length"), n, None if dst_base is None else self._check_dev(dst_base, 1, "dst_base"),
            dst_off.data_ptr(), total.data_ptr(), stream)
        N.check_call(rc, "snp_pack_batch")
        return dst_off, total

    def find_match_length_batch_device(self, base, s1, s2, s2_limit, stream: int | None = None):
        """Batched SnappyCompressor.FindMatchLength on the device (u8 base, i32 positions); returns i32[N] on the device."""
        import torch
        n = s1.numel()
        out = torch.zeros(n, dtype=torch.int32, device=s1.device)
        rc = N.lib().snp_find_match_length_batch(
            self._ctx, self._check_dev(base, 1, "base"), self._check_dev(s1, 4, "s1"), self._check_dev(s2, 4, "s2"),
            self._check_dev(s2_limit, 4, "s2_limit"), out.data_ptr(), n, stream)
        N.check_call(rc, "snp_find_match_length_batch")
        return out

    # --------------------------------------------------------------- host mode
    def compress_batch_host(self, in_base: np.ndarray, in_off, in_len, out_base: np.ndarray, out_off,
                            out_cap, hash_mode: int = N.HASH_CRC32C):
        """Host buffers in, host buffers out (PCIe both ways inside the call).
        Returns (out_len u32[N], status i32[N])."""
        in_off = np.ascontiguousarray(in_off, np.uint64)
        in_len = np.ascontiguousarray(in_len, np.uint32)
        out_off = np.ascontiguousarray(out_off, np.uint64)
        out_cap = np.ascontiguousarray(out_cap, np.uint32)
        n = in_off.size
        out_len = np.zeros(n, np.uint32)
        status = np.zeros(n, np.int32)
        rc = N.lib().snp_compress_batch(self._ctx, _np_ptr(in_base), _np_ptr(in_off), _np_ptr(in_len),
                                        _np_ptr(out_base), _np_ptr(out_off), _np_ptr(out_cap),
                                        _np_ptr(out_len), _np_ptr(status), n, hash_mode, N.MEM_HOST, None)
        N.check_call(rc, "snp_compress_batch")
        return out_len, status

    def decompress_batch_host(self, in_base: np.ndarray, in_off, in_len, out_base: np.ndarray, out_off,
                              out_cap):
        in_off = np.ascontiguousarray(in_off, np.uint64)
        in_len = np.ascontiguousarray(in_len, np.uint32)
        out_off = np.ascontiguousarray(out_off, np.uint64)
        out_cap = np.ascontiguousarray(out_cap, np.uint32)
        n = in_off.size
        out_len = np.zeros(n, np.uint32)
        status = np.zeros(n, np.int32)
        rc = N.lib().snp_decompress_batch(self._ctx, _np_ptr(in_base), _np_ptr(in_off), _np_ptr(in_len),
                                          _np_ptr(out_base), _np_ptr(out_off), _np_ptr(out_cap),
                                          _np_ptr(out_len), _np_ptr(status), n, N.MEM_HOST, None)
        N.check_call(rc, "snp_decompress_batch")
        return out_len, status

    def uncompressed_length_batch_host(self, in_base: np.ndarray, in_off, in_len):
        in_off = np.ascontiguousarray(in_off, np.uint64)
        in_len = np.ascontiguousarray(in_len, np.uint32)
        n = in_off.size
        ulen = np.zeros(n, np.uint32)
        status = np.zeros(n, np.int32)
        rc = N.lib().snp_uncompressed_length_batch(self._ctx, _np_ptr(in_base), _np_ptr(in_off),
                                                   _np_ptr(in_len), _np_ptr(ulen), _np_ptr(status), n,
                                                   N.MEM_HOST, None)
        N.check_call(rc, "snp_uncompressed_length_batch")
        return ulen, status


# ---- list-of-bytes conveniences used by tests and smoke() --------------------

def pack(items) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Concatenate byte strings -> (base u8, off u64[N], len u32[N])."""
    lens = np.array([len(x) for x in items], np.uint32)
    offs = np.zeros(len(items), np.uint64)
    if len(items):
        offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    base = np.frombuffer(b"".join(bytes(x) for x in items), np.uint8).copy() if len(items) else np.zeros(0, np.uint8)
    if base.size == 0:
        base = np.zeros(1, np.uint8)
    return base, offs, lens


def compress_many(engine: Engine, items, hash_mode: int = N.HASH_CRC32C):
    """Compress each byte string (<= 64 KiB) as its own block -> (list[bytes], status[N])."""
    base, offs, lens = pack(items)
    caps = np.array([N.lib().snp_get_max_compressed_length(int(l)) for l in lens], np.uint32)
    out_off = np.zeros(len(items), np.uint64)
    if len(items):
        out_off[1:] = np.cumsum(caps[:-1], dtype=np.uint64)
    out = np.zeros(int(caps.sum()) + 1, np.uint8)
    out_len, status = engine.compress_batch_host(base, offs, lens, out, out_off, caps, hash_mode)
    res = [out[int(o): int(o) + int(l)].tobytes() for o, l in zip(out_off, out_len)]
    return res, status


def decompress_many(engine: Engine, items, caps=None):
    """Decompress each block -> (list[bytes], status[N]).  caps default to the declared lengths."""
    base, offs, lens = pack(items)
    if caps is None:
        ulen, _ = engine.uncompressed_length_batch_host(base, offs, lens)
        caps = ulen
    caps = np.ascontiguousarray(caps, np.uint32)
    out_off = np.zeros(len(items), np.uint64)
    if len(items):
        out_off[1:] = np.cumsum(caps[:-1], dtype=np.uint64)
    out = np.zeros(int(caps.astype(np.uint64).sum()) + 1, np.uint8)
    out_len, status = engine.decompress_batch_host(base, offs, lens, out, out_off, caps)
    res = [out[int(o): int(o) + int(l)].tobytes() for o, l in zip(out_off, out_len)]
    return res, status
