"""CPU tests: the C-ABI shared library builds, loads and exports every symbol
include/snappier_b200.h declares; sizing works; compute fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from snappier_b200 import _native as N, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "snappier_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(snp_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(N.SYMBOLS), (declared, sorted(N.SYMBOLS))
    L = N.lib()
    for s in declared:
        assert getattr(L, s) is not None
    assert L.snp_abi_version() == 1


def test_status_codes_match_oracle(oracle):
    from snappier_b200 import _native as N
    for name in ("OK", "OUTPUT_TOO_SMALL", "INVALID_LENGTH", "INCOMPLETE", "INVALID_COPY_OFFSET", "DATA_TOO_LONG"):
        assert getattr(N, name) == getattr(oracle, name)
    assert N.status_string(N.INCOMPLETE) == "Incomplete Snappy block."
    assert N.status_string(N.INVALID_COPY_OFFSET) == "Invalid copy offset"
    assert N.status_string(N.DATA_TOO_LONG) == "Data too long"
    assert N.status_string(N.OUTPUT_TOO_SMALL) == "Output buffer is too small."


def test_sizing_and_host_varint(oracle, kats):
    """Snappy.GetMaxCompressedLength (Snappy.cs:20-24) and GetUncompressedLength (:142-143)
    need no GPU."""
    from snappier_b200 import snappy as S
    for n in (0, 1, 100, 65535, 65536, 65537, 1 << 20, 100000):
        assert S.get_max_compressed_length(n) == oracle.get_max_compressed_length(n) == 32 + n + n // 6 + 1 + 5
    for k in kats["varint"]:
        if k["value"] <= 0x7fffffff:
            assert S.get_uncompressed_length(bytes(k["bytes"])) == k["value"]
    for bad in kats["varint_incomplete"] + [[0xff] * 6, []]:
        with pytest.raises(S.InvalidDataException):
            S.get_uncompressed_length(bytes(bad))


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    from snappier_b200 import _native as N, snappy as S
    from snappier_b200.batch import Engine
    with pytest.raises(N.NativeLibraryError):
        Engine(0)
    with pytest.raises(N.NativeLibraryError):
        S.compress_to_array(b"hello hello hello hello")
    with pytest.raises(N.NativeLibraryError):
        S.decompress_to_array(b"\x05\x10hello")
    out = np.zeros(8, np.uint8)
    w = C.c_size_t(0)
    assert N.lib().snp_compress(None, 0, out.ctypes.data, 8, C.byref(w), 0) == N.E_NO_DEVICE


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under snappier_b200/ or include/ may reference it."""
    for base in ("snappier_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    for pat in (r'#\s*include\s*[<"][^>"]*oracle', r'^\s*(from|import)\s+oracle', r'\borc_\w+\s*\(',
                                r'libsnappy_oracle', r'pyoracle', r'dlopen[^\n]*oracle'):
                        assert not re.search(pat, txt, re.M), (f, pat)  # doc comments may cite oracle/ by path


def test_numa_helper_parses_cpulists_and_is_a_noop_without_sysfs():
    """snappier_b200.numa: cpulist parsing, and binding degrades to a no-op where the GPU's node is unknown."""
    from snappier_b200 import numa
    assert numa._parse_cpulist("0-3,8,10-11") == [0, 1, 2, 3, 8, 10, 11]
    assert numa._parse_cpulist("5") == [5]
    info = numa.bind_to_gpu_node(0)  # no GPU here: nothing may change
    assert info["bound"] is False and info["cpus"] >= 1
