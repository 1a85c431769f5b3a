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
def test_cpp_mirror_kats(tmp_path, oracle):
    import numpy as np

    exe = _build(tmp_path)
    out = str(tmp_path)
    r = subprocess.run([exe, out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cpp mirror ok" in r.stdout, r.stdout + r.stderr
    # the device-resident graphs the program ran, against the CPU oracle (bit patterns)
    x = np.fromfile(os.path.join(out, "chain_x.bin"), np.float32)
    y = np.fromfile(os.path.join(out, "chain_y.bin"), np.float32)
    ba = np.fromfile(os.path.join(out, "chain_ba.bin"), np.float32)
    stw = np.fromfile(os.path.join(out, "chain_state.bin"), np.float32)
    lanes, k = 48, 4
    so = np.zeros((stw.size // lanes, lanes), np.float32)
    want = oracle.chain_lanes(k, ba, so, x, lanes, 1)
    assert np.array_equal(y.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(stw.view(np.uint32), so.reshape(-1).view(np.uint32))
    xl = np.fromfile(os.path.join(out, "lockin_x.bin"), np.int32)
    step = np.fromfile(os.path.join(out, "lockin_step.bin"), np.int32)
    iq = np.fromfile(os.path.join(out, "lockin_iq.bin"), np.int32)
    lanes = step.size
    want = oracle.lockin_lanes([1048576, -94906265], np.zeros(lanes, np.int32), step, np.zeros((4, lanes), np.int64), xl, lanes, 0)
    assert np.array_equal(iq, want)
