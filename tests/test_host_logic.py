"""Host-side mirror logic that needs no GPU: views, Split/Lanes plumbing, Q, Filter."""
import numpy as np
import pytest

from idsp_b200 import (Biquad, BiquadClamp, DirectForm1, Filter, Lanes, Q8, Q32, Split, View,
                       hbf_dec_response_length, hbf_int_response_length, hbf_taps)
from idsp_b200.process import FRAME_MAJOR, LANE_MAJOR


def test_view_from_flat_asserts_length():
    """dsp-process/src/view.rs:181-182"""
    with pytest.raises(AssertionError):
        View.from_flat(np.zeros(7, np.int32), 3, 2)
    v = View.from_flat(np.arange(6, dtype=np.int32), 3, 2)
    assert v.layout == LANE_MAJOR and list(v.lane(1)) == [3, 4, 5]
    f = View.from_frames(np.arange(6, dtype=np.int32), 2)
    assert f.layout == FRAME_MAJOR and f.frames == 3 and list(f.frame(1)) == [2, 3]


def test_split_lanes_replicates_state():
    """dsp-process/src/split.rs:272-277"""
    st = DirectForm1.default("i32", 1)
    st.words[:, 0] = [1, 2, 3, 4]
    p = Split.new(Biquad.identity(Q32(30)), st).lanes(5)
    assert isinstance(p.config, Lanes) and p.state.lanes == 5
    assert np.array_equal(p.state.words[:, 3], [1, 2, 3, 4])


def test_q_rounding_and_saturation():
    assert Q32(30).from_float(2.0) == (1 << 31) - 1  # saturating `as`
    assert Q32(30).from_float(-2.0) == -(1 << 31)
    assert Q8(4).from_float(1.5) == 24
    assert Q8(0).from_float(0.5) == 1 and Q8(0).from_float(-0.5) == -1  # half away from zero
    assert Q8(0).from_float(0.49999999999999994) == 0
    assert Q32(0).from_float(float("nan")) == 0


def test_biquad_constants_and_clamp_defaults():
    b = Biquad.identity(Q32(30))
    assert list(b.ba) == [1 << 30, 0, 0, 0, 0]
    assert list(Biquad.hold("f32").ba) == [0, 0, 0, 1, 0]
    assert Biquad.proportional(3.0, "f32").forward_gain() == 3.0
    c = BiquadClamp(Biquad.proportional(3.0, "f32"))
    c.u = 6.0
    assert c.input_offset() == 2.0
    c.set_input_offset(2.0)
    assert c.u == 6.0
    ci = BiquadClamp(Biquad.identity(Q32(30)))
    assert (ci.u, ci.min, ci.max) == (0, -(1 << 31), (1 << 31) - 1)


def test_filter_builder():
    ba = Filter().critical_frequency(0.1).set_gain(1000.0).lowpass()
    assert list(Biquad.from_ba6(ba, Q32(30)).ba) == [2147483647, 2147483647, 2147483647, 1227265970, -443242341]


def test_hbf_constants():
    t = hbf_taps()
    assert [len(x) for x in t] == [23, 10, 5, 4, 3]
    assert hbf_dec_response_length(4) == 57 and hbf_int_response_length(4) == 922


def test_block_length_mismatch_is_an_error():
    """the reference debug_asserts equal lengths (process.rs:121-123)"""
    st = DirectForm1.default("i32", 2)
    with pytest.raises(ValueError):
        Biquad.identity(Q32(30)).block(st, np.zeros(4, np.int32), np.zeros(6, np.int32))
    with pytest.raises(ValueError):
        Biquad.identity(Q32(30)).block(st, np.zeros(3, np.int32), np.zeros(3, np.int32))


def test_peer_view_is_a_device_output():
    """a range of another process's buffer (idsp_b200_ipc_open) is passed to the kernels by address"""
    import torch

    from idsp_b200.engine import PeerView, _is_dev, _ptr

    v = PeerView(0x7F0000001000, 64, torch.int32)
    assert _is_dev(v) and v.numel() == 64 and _ptr(v).value == 0x7F0000001000
