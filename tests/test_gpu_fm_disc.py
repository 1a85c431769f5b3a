"""GPU parity of the fused FM-discriminator graph (examples/fm_disc.rs:26-48)
against the CPU oracle, bit-exact, plus the example's own self-test run on the GPU."""
import numpy as np
import pytest
import torch

from gpu_common import DEV, assert_bits_equal, to_dev, to_np

pytestmark = pytest.mark.gpu

from idsp_b200 import Biquad, Filter, FmDiscriminator, FmDiscState, Lanes, Q32
from test_oracle_fm_disc import _lowpass_f32, run_fm_disc, TAU


def _flat(x_tlc, layout):  # [frames, lanes, 2] -> flat pairs in the layout
    return np.ascontiguousarray(x_tlc if layout == 0 else np.swapaxes(x_tlc, 0, 1)).reshape(-1)


def test_tracks_known_modulation_on_gpu(oracle):
    """examples/fm_disc.rs:149-158"""
    _, m, x, ba = run_fm_disc(oracle)
    proc = Lanes(FmDiscriminator(0x19341234, Biquad(ba, Q32(30))))
    st = FmDiscState.default(1, DEV)
    y = torch.empty(len(x), dtype=torch.int32, device=DEV)
    proc.block(st, to_dev(x.reshape(-1)), y)
    y = (to_np(y).astype(np.float32) * (TAU / np.float32(4294967296.0)))[1024:].astype(np.float64)
    m = m.astype(np.float64)
    gain = (y * m).sum() / (m * m).sum()
    assert (y * m).sum() / np.sqrt((y * y).sum() * (m * m).sum()) > 0.999
    assert 0.95 < gain < 1.05 and np.sqrt(((y - gain * m) ** 2).sum()) / len(y) < 5e-4


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("lanes,frames", [(1, 257), (33, 100), (300, 64), (256, 96), (1028, 72)])
@pytest.mark.parametrize("policy", [0, 1])
def test_fm_disc_vs_oracle(oracle, layout, lanes, frames, policy):
    """policy 0: TMA kernels where the chunk qualifies (8-byte in, 4-byte out), policy 1: generic kernels"""
    import idsp_b200 as ib
    ib.default_context(0).set_kernel_policy(policy)
    try:
        _fm_disc_vs_oracle(oracle, layout, lanes, frames)
    finally:
        ib.default_context(0).set_kernel_policy(0)


def _fm_disc_vs_oracle(oracle, layout, lanes, frames):
    rng = np.random.default_rng(lanes * 7 + frames)
    ph = rng.integers(-(1 << 31), 1 << 31, (frames, lanes)).astype(np.int32)
    x = oracle.cossin(ph.reshape(-1)).reshape(frames, lanes, 2)
    x[:, ::4] = rng.integers(-(1 << 31), 1 << 31, x[:, ::4].shape).astype(np.int32)  # arbitrary bits incl. extremes
    x[0, 0] = (-(1 << 31), -(1 << 31))
    ba = Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), Q32(30)).ba
    so = np.zeros((7, lanes), np.int32)
    st = FmDiscState.default(lanes, DEV)
    proc = Lanes(FmDiscriminator(0x9abc1234, Biquad(ba, Q32(30))))
    a = 0
    for c in (frames // 2, 1, frames - frames // 2 - 1):
        xc = _flat(x[a:a + c], layout)
        want = oracle.fm_disc_lanes(0x9abc1234, ba, 30, so, xc, lanes, layout)
        y = torch.empty(c * lanes, dtype=torch.int32, device=DEV)
        proc.block(st, to_dev(xc), y, layout)
        assert_bits_equal(to_np(y), want, f"fm_disc lanes={lanes} layout={layout}")
        a += c
    assert_bits_equal(st.numpy(), so, "state")
