"""snappier_b200 -- B200-native Snappy block engine behind Snappier's block API.

Layout (only what the hot path needs):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/snappier_b200.h)
  _native.py       ctypes loader of libsnappier_b200.so (fails loudly, no fallback)
  snappy.py        mirror of Snappier's `Snappy` static facade (single-call API)
  batch.py         batched API (device tensors or host arrays)
  sharding.py      block-range sharding across the GPUs of one box
  build.py         nvcc build of the shared library
"""
from . import _native  # noqa: F401
from ._native import (BLOCK_SIZE, HASH_CRC32C, HASH_MUL, NativeLibraryError)  # noqa: F401

__all__ = ["BLOCK_SIZE", "HASH_CRC32C", "HASH_MUL", "NativeLibraryError"]
