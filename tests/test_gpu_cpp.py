"""Compiled host side: builds tests/cpp/test_mirror.cpp (C++17 mirror of the reference's operator
surface, include/idsp_b200.hpp) against libidsp_b200.so and runs the reference's KATs through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = os.path.join(str(tmp_path), "test_mirror")
    lib = os.path.join(ROOT, "idsp_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"), "-L", lib, "-lidsp_b200",
                    f"-Wl,-rpath,{lib}", "-o", exe], check=True)
    return exe


def test_cpp_mirror_compiles(tmp_path):
    """host-only: header + test program compile and link against the C ABI (no GPU needed)"""
    from idsp_b200.build import build

    build()
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_cpp_mirror_kats(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cpp mirror ok" in r.stdout, r.stdout + r.stderr
