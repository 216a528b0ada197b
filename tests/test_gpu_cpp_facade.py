"""Builds tests/cpp/test_snappy_facade.cpp (g++) against include/snappier_b200.hpp and runs it
on the GPU: the C++ host-side mirror of Snappier's `Snappy` facade, checked against oracle bytes."""
import os
import subprocess

import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp) -> str:
    exe = os.path.join(tmp, "test_snappy_facade")
    pkg = os.path.join(ROOT, "snappier_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_snappy_facade.cpp"), "-o", exe,
                           "-L", pkg, "-l:libsnappier_b200.so", f"-Wl,-rpath,{pkg}"])
    return exe


def test_cpp_facade_compiles(tmp_path):
    """CPU check: header + test program compile and link against the C-ABI library."""
    from snappier_b200 import build
    build.build()
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
def test_cpp_facade_runs_like_reference_tests(tmp_path, oracle, fixtures):
    exe = _build(str(tmp_path))
    data = fixtures["corpus/html"] + fixtures["corpus/alice29.txt"][:50000]
    open(tmp_path / "input.bin", "wb").write(data)
    open(tmp_path / "input.snappy", "wb").write(oracle.compress(data)[1])
    open(tmp_path / "golden_framed.snappy", "wb").write(fixtures["framed/html_x_4.snappy"])
    open(tmp_path / "golden_raw.bin", "wb").write(fixtures["corpus/html_x_4"])
    for i in (1, 2, 3):
        open(tmp_path / f"baddata{i}.snappy", "wb").write(fixtures[f"bad/baddata{i}.snappy"])
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
