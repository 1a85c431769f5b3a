"""Host-side coefficient builders of the C ABI (no GPU needed): `idsp_filter_*`,
`idsp_biquad_from_*`, `idsp_pid_*` against the reference's own doctests / unit tests and against the
independent Python restatement in idsp_b200/coefficients.py."""
import ctypes as C
import math

import numpy as np
import pytest

import idsp_b200 as ib
from idsp_b200 import Filter, FilterError, Q32, _lib, pid
from idsp_b200.coefficients import TYPES, WebAudio
from idsp_b200.iir import Biquad, Q


def test_lowpass_highpass_golden_coefficients():
    """src/iir/coefficients.rs:289-301, 316-327: f0 = 0.1, gain = 1000 -> Biquad<Q32<30>>; the three b taps
    saturate at i32::MAX (highpass b1 at i32::MIN); a1 / a2 as verified in SURVEY 8(c)"""
    f = Filter().critical_frequency(0.1).set_gain(1000.0)
    assert f.build_biquad("lowpass", Q32(30)).ba.tolist() == [2147483647, 2147483647, 2147483647, 1227265970, -443242341]
    assert f.build_biquad("highpass", Q32(30)).ba.tolist() == [2147483647, -2147483648, 2147483647, 1227265970, -443242341]
    # the one-call C entry point
    st = _lib.FilterF64()
    _lib.lib().idsp_filter_default_f64(C.byref(st))
    st.frequency, st.gain = math.tau * 0.1, 1000.0
    out = (C.c_int32 * 5)()
    assert _lib.lib().idsp_filter_build_biquad_f64(C.byref(st), 0, _lib.KIND_CODE["i32"], 30, out) == 0
    assert list(out) == [2147483647, 2147483647, 2147483647, 1227265970, -443242341]


def test_filter_default_matches_reference():
    """coefficients.rs:88-97 and :18-22"""
    st = _lib.FilterF64()
    _lib.lib().idsp_filter_default_f64(C.byref(st))
    assert (st.frequency, st.gain, st.shelf, st.shape_kind) == (0.0, 1.0, 1.0, 0)
    assert st.shape == 1.0 / math.sqrt(2.0)
    s32 = _lib.FilterF32()
    _lib.lib().idsp_filter_default_f32(C.byref(s32))
    assert np.float32(s32.shape) == np.float32(1.0) / np.float32(math.sqrt(2.0))


@pytest.mark.parametrize("typ", TYPES)
def test_c_builder_equals_python_restatement(typ):
    """two restatements of coefficients.rs:302-464 written separately must agree bit for bit in f64"""
    rng = np.random.default_rng(TYPES.index(typ))
    for _ in range(200):
        f = Filter().angular_critical_frequency(float(rng.uniform(1e-4, math.pi - 1e-4)))
        f.set_gain(float(10 ** rng.uniform(-3, 3))).set_shelf(float(10 ** rng.uniform(-2, 2)))
        sel = rng.integers(3)
        if sel == 0:
            f.q(float(10 ** rng.uniform(-1, 1.5)))
        elif sel == 1:
            f.bandwidth(float(rng.uniform(0.1, 4)))
        else:
            f.shelf_slope(float(rng.uniform(0.1, 1.0)))
        a, b = np.array(getattr(f, typ)()), np.array(f.build_c(typ))
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (typ, f, a, b)


def test_f32_flavour_rounds_in_f32():
    """`Filter<f32>` (examples/fm_disc.rs: Filter<f32> -> Biquad<Q32<30>>): every intermediate is an f32, so the
    result is exactly representable in f32 and close to, but not the same as, the f64 result rounded"""
    f64 = Filter().critical_frequency(0.02)
    f32 = Filter(dtype="f32").critical_frequency(0.02)
    a, b = np.array(f64.build("lowpass")), np.array(f32.build("lowpass"))
    assert np.array_equal(b, b.astype(np.float32).astype(np.float64))
    assert np.allclose(a, b, rtol=1e-5, atol=1e-9)
    q64 = f64.build_biquad("lowpass", Q32(30)).ba
    q32 = f32.build_biquad("lowpass", Q32(30)).ba
    assert np.max(np.abs(q64.astype(np.int64) - q32.astype(np.int64))) < 1 << 8  # a few f32 ULPs of 2^30
    # f32 quantisation: (v * 2^30).round() computed in f32 -> multiples of 64 around 2^30
    assert all(int(v) % 64 == 0 for v in q32[3:4])


def test_validate_messages():
    """coefficients.rs:241-265: the iir::Error variant and field"""
    cases = [
        (dict(frequency=float("nan")), "NonFinite(frequency)"),
        (dict(frequency=-0.1), "OutOfRange(frequency)"),
        (dict(frequency=3.2), "OutOfRange(frequency)"),
        (dict(gain=0.0), "NonPositive(gain)"),
        (dict(gain=float("inf")), "NonPositive(gain)"),
        (dict(shelf=-1.0), "NonPositive(shelf)"),
    ]
    L = _lib.lib()
    for kw, msg in cases:
        f = Filter(**kw)
        with pytest.raises(FilterError, match=msg.replace("(", r"\(").replace(")", r"\)")):
            f.validate()
        assert L.idsp_filter_validate_f64(C.byref(f._c_struct())) == -1
        assert L.idsp_b200_last_error().decode() == msg
    for shape, msg in ((("q", float("nan")), "NonFinite(q)"), (("q", 0.0), "NonPositive(q)"),
                       (("bandwidth", float("inf")), "NonFinite(bandwidth)"), (("slope", -1.0), "NonPositive(slope)")):
        f = Filter()
        f.shape = ib.coefficients.Shape(*shape)
        assert L.idsp_filter_validate_f64(C.byref(f._c_struct())) == -1
        assert L.idsp_b200_last_error().decode() == msg
        with pytest.raises(FilterError):
            f.validate()
    f = Filter().critical_frequency(0.25).bandwidth(-2.0)  # negative bandwidth is accepted by the reference
    f.validate()
    assert L.idsp_filter_validate_f64(C.byref(f._c_struct())) == 0
    with pytest.raises(FilterError):
        Filter(frequency=4.0).try_build("lowpass")


def test_quantisation_rules():
    """(v * 2^F).round() as T: half away from zero, saturating, NaN -> 0 (num_traits_impl.rs:32-45 + Rust `as`);
    C builder == Python Q.from_float for every kind"""
    L = _lib.lib()
    vals = [0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, 0.49999999999999994, 1e30, -1e30, float("nan"), float("inf"), -float("inf"),
            0.999999, -1.0, 127.5 / 128, 1.0]
    for kind, F in (("i8", 7), ("i8", 0), ("i16", 14), ("i32", 30), ("i32", 0), ("i64", 62), ("i32", -2), ("i16", 20)):
        for v in vals:
            out = np.zeros(5, {"i8": np.int8, "i16": np.int16, "i32": np.int32, "i64": np.int64}[kind])
            ba5 = (C.c_double * 5)(v, v, v, v, v)
            assert L.idsp_biquad_from_ba5_f64(ba5, _lib.KIND_CODE[kind], F, out.ctypes.data_as(C.c_void_p)) == 0
            assert int(out[0]) == Q(kind, F).from_float(v), (kind, F, v, out[0])
    # f32 -> Q32<32> as used by PLL::from_zpk (src/pll.rs:41-46)
    for v in (0.1, -0.3, 0.49999997, 0.5, -0.5, 1e-9):
        out = np.zeros(5, np.int32)
        ba5 = (C.c_float * 5)(*[v] * 5)
        assert L.idsp_biquad_from_ba5_f32(ba5, _lib.KIND_CODE["i32"], 32, out.ctypes.data_as(C.c_void_p)) == 0
        assert int(out[0]) == ib.PLL._q32(np.float32(v))


def test_from_zpk():
    """biquad.rs:578-619: [gain, -gain*(z0+z1), gain*z0*z1, p0+p1, -p0*p1]"""
    L = _lib.lib()
    out = np.zeros(5, np.float64)
    z, p = (C.c_double * 2)(0.5, -0.25), (C.c_double * 2)(0.9, 0.1)
    assert L.idsp_biquad_from_zpk_f64(z, 0, p, 1, 2.0, _lib.KIND_CODE["f64"], 0, out.ctypes.data_as(C.c_void_p)) == 0
    assert out.tolist() == [2.0, -2.0 * (0.5 + -0.25), 2.0 * (0.5 * -0.25), 0.9 + 0.9, -(0.9 * 0.9 + 0.1 * 0.1)]


def test_pid_reference_vectors():
    """src/iir/pid.rs:573-590 (`fn pid`), :253-259 (proportional), :605-619 (`fn units`), :96-103 (I gain doctest)"""
    b = pid.Builder(dtype="f32").gain("I", 1e-3).gain("P", 1.0).gain("D", 1e2).limit("I", 1e3).limit("D", 1e1).build(1.0, "f32")
    want = np.array([9.181_909, -18.272_726, 9.090_908, 1.909_090_8, -0.909_090_8], np.float32)
    assert np.all(np.abs(b.ba / want - 1.0) < 2.0 * np.finfo(np.float32).eps), b.ba
    p = pid.Builder(dtype="f32").gain("P", 3.0).set_order("P").build(1.0, "f32")
    assert p.ba.tolist() == [3.0, 0.0, 0.0, 0.0, 0.0]  # == Biquad::proportional(3.0)
    # units(): an integrator built for period tau steps by tau * ki per sample
    ki, tau = 5e-2, 3e-3
    i = pid.Builder(dtype="f32").gain("I", ki).build(tau, "f32")
    x1 = x2 = y1 = y2 = np.float32(0)
    for n in range(1, 10):
        x0 = np.float32(1.0)
        y0 = np.float32(np.float32(np.float32(np.float32(i.ba[0] * x0 + i.ba[1] * x1) + i.ba[2] * x2) + i.ba[3] * y1) + i.ba[4] * y2)
        x2, x1, y2, y1 = x1, x0, y1, y0
        assert abs(y0 / (np.float32(n) * np.float32(tau) * np.float32(ki)) - 1.0) < 3.0 * np.finfo(np.float32).eps
    # Q29 flavour (`fn pid_i32`): the GAINS are quantised, then accumulated in the integer type
    q = pid.Builder(dtype="f32").gain("I", 1e-5).gain("P", 1e-2).gain("D", 1e0).limit("I", 1e1).limit("D", 1e-1).build(1.0, Q32(29))
    f = pid.Builder(dtype="f32").gain("I", 1e-5).gain("P", 1e-2).gain("D", 1e0).limit("I", 1e1).limit("D", 1e-1).build(1.0, "f64")
    assert np.max(np.abs(q.ba / 2.0 ** 29 - f.ba)) < 8 / 2.0 ** 29
    assert abs(int(q.ba[0]) + int(q.ba[1]) + int(q.ba[2])) < 1 << 20  # forward DC gain ~ ki * a0i


def test_pid_validate():
    """pid.rs:193-222"""
    b = pid.Builder().gain("I", 1.0)
    b.validate(1e-3)
    for period, msg in ((float("nan"), "NonFinite(period)"), (0.0, "NonPositive(period)"), (-1.0, "NonPositive(period)")):
        with pytest.raises(pid.PidError, match=msg.replace("(", r"\(").replace(")", r"\)")):
            b.validate(period)
    with pytest.raises(pid.PidError, match="SignMismatch"):
        pid.Builder().gain("I", 1.0).limit("I", -5.0).validate(1.0)
    with pytest.raises(pid.PidError, match=r"NonPositive\(limit\)"):
        pid.Builder().gain("D", 1.0).limit("D", 0.0).validate(1.0)
    with pytest.raises(pid.PidError, match=r"NonFinite\(gain\)"):
        pid.Builder().gain("P", float("nan")).validate(1.0)
    with pytest.raises(pid.PidError):
        pid.Builder().gain("I", 1.0).limit("I", -5.0).try_build(1.0, "f32")


def test_webaudio_defaults():
    """coefficients.rs:99-109, 529-543"""
    w = WebAudio()
    f = w.filter()
    assert f.frequency == math.tau * (350.0 / 48e3) and f.shape.value == 1.0 and f.shelf == 1.0
    w = WebAudio(typ="peaking", gain_db=6.0)
    assert abs(w.filter().shelf - 10 ** 0.3) < 1e-12
    assert len(w.build()) == 2


def test_build_clamped_defaults():
    b = Filter().critical_frequency(0.1).build_clamped("lowpass", "f32")
    assert b.u == 0 and b.min == -np.inf and b.max == np.inf
    b = Filter().critical_frequency(0.1).build_clamped("lowpass", Q32(30))
    assert (b.min, b.max) == (-(1 << 31), (1 << 31) - 1)
