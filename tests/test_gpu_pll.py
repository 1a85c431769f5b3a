"""GPU parity of the PLL lanes against the CPU oracle (bit-exact wrapping
integer math) and the reference's convergence tests run on the GPU (src/pll.rs:117-149)."""
import numpy as np
import pytest
import torch

from gpu_common import DEV, assert_bits_equal, layout_flat, to_dev, to_np

pytestmark = pytest.mark.gpu

from idsp_b200 import PLL, Lanes, PLLState


def _ramp(step, n, start=0):
    k = np.arange(start + 1, start + n + 1, dtype=np.uint64)
    return ((k * np.uint64(step)) & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32)


def _wrap32(v):
    return (int(v) + (1 << 31)) % (1 << 32) - (1 << 31)


def test_host_coefficients_match_oracle(oracle):
    for bw in (5e-2, 8e-5, 1e-3):
        assert PLL.from_bandwidth(bw, 4.0).ba == [int(v) for v in oracle.pll_from_bandwidth(bw, 4.0)]


def test_converge_pll_on_gpu():
    """src/pll.rs:117-132"""
    p = PLL.from_bandwidth(5e-2, 4.0)
    n, step = 1 << 9, 0x71f63049
    x = _ramp(step, n)
    st = PLLState.default(1, DEV)
    half = n // 2 + 1
    y = torch.empty(half, dtype=torch.int32, device=DEV)
    Lanes(p).block(st, to_dev(x[:half]), y)
    for i in range(half, n):
        yi = torch.empty(1, dtype=torch.int32, device=DEV)
        Lanes(p).block(st, to_dev(x[i:i + 1]), yi)
        assert abs(_wrap32(step + int(st.frequency()[0]))) <= 1
        assert abs(_wrap32(int(x[i]) + int(to_np(yi)[0]))) <= 4


def test_converge_narrow_on_gpu():
    """src/pll.rs:134-149"""
    p = PLL.from_bandwidth(8e-5, 4.0)
    n, step = 1 << 18, 0x1401235
    x = _ramp(step, n)
    st = PLLState.default(1, DEV)
    half = n // 2 + 1
    y = torch.empty(n, dtype=torch.int32, device=DEV)
    Lanes(p).block(st, to_dev(x[:half]), y[:half])
    Lanes(p).block(st, to_dev(x[half:]), y[half:])
    assert abs(_wrap32(step + int(st.frequency()[0]))) <= 1 << 16
    e = (x[half:].astype(np.int64) + to_np(y)[half:].astype(np.int64) + (1 << 31)) % (1 << 32) - (1 << 31)
    assert np.all(np.abs(e) <= 1 << 16)


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("lanes,frames", [(1, 300), (37, 211), (256, 128), (1000, 64), (64, 132)])
def test_pll_vs_oracle(oracle, layout, lanes, frames):
    rng = np.random.default_rng(lanes + frames)
    ba = oracle.pll_from_bandwidth(3e-3, 4.0)
    steps = rng.integers(1, 1 << 32, lanes, dtype=np.uint64)
    x = np.stack([_ramp(int(s), frames) for s in steps], axis=1)  # [frames, lanes] phase ramps
    x[:, ::3] = rng.integers(-(1 << 31), 1 << 31, (frames, x[:, ::3].shape[1])).astype(np.int32)  # noise lanes: clamp engages
    so = np.zeros((9, lanes), np.int32)
    st = PLLState.default(lanes, DEV)
    a = 0
    for c in (frames // 3, 1, frames - frames // 3 - 1):  # streaming: state carried across calls
        xc = layout_flat(x[a:a + c], layout)
        want = oracle.pll_lanes(ba, so, xc, lanes, layout)
        y = torch.empty(c * lanes, dtype=torch.int32, device=DEV)
        Lanes(PLL(ba)).block(st, to_dev(xc), y, layout)
        assert_bits_equal(to_np(y), want, f"pll lanes={lanes} layout={layout}")
        a += c
    assert_bits_equal(st.numpy(), so, "state")


def test_pll_large_lock_property():
    """bench-like size: every lane locks to its own ramp (frequency estimate == -step within 1 LSB... 2^8)"""
    lanes, frames = 16384, 4096
    rng = np.random.default_rng(3)
    steps = rng.integers(1 << 20, 1 << 31, lanes, dtype=np.uint64)
    k = torch.arange(1, frames + 1, dtype=torch.int64, device=DEV)[:, None]
    x = ((k * torch.from_numpy(steps.astype(np.int64)).to(DEV)[None, :]) & 0xffffffff)
    x = torch.where(x >= (1 << 31), x - (1 << 32), x).to(torch.int32).reshape(-1)
    st = PLLState.default(lanes, DEV)
    y = torch.empty_like(x)
    Lanes(PLL.from_bandwidth(2e-2, 4.0)).block(st, x, y)
    err = (steps.astype(np.int64) + st.frequency().astype(np.int64) + (1 << 31)) % (1 << 32) - (1 << 31)
    assert np.all(np.abs(err) <= 1 << 8)
