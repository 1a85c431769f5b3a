"""GPU parity of the iir::Biquad family against the CPU oracle (bit-exact, through the C ABI)."""
import numpy as np
import pytest
import torch

from gpu_common import BITS, DEV, NP, assert_bits_equal, layout_flat, rand_samples, to_dev, to_np

pytestmark = pytest.mark.gpu

import idsp_b200 as ib
from idsp_b200 import (Biquad, BiquadClamp, Cascade, DirectForm, DirectForm1, DirectForm1Dither,
                       DirectForm1Wide, DirectForm2Transposed, Filter, Lanes, Q32, Split, View)
from idsp_b200.iir import Q


def _coeffs(kind, rng):
    if kind in BITS:
        F = BITS[kind] - 2
        ba6 = Filter().critical_frequency(0.05).lowpass()
        return Biquad.from_ba6(ba6, Q(kind, F))
    return Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), kind)


# ------------------------------------------------------------------ reference KATs on the GPU
def test_kat_lowpass_highpass_golden():
    """src/iir/coefficients.rs:289-301, 316-327 through the GPU"""
    for ba6, want in ((Filter().critical_frequency(0.1).set_gain(1000.0).lowpass(), [5, 3, 9, 25, 42, 49]),
                      (Filter().critical_frequency(0.1).set_gain(1000.0).highpass(), [5, -9, 11, 12, -1, 17])):
        iir = Biquad.from_ba6(ba6, Q32(30))
        xy = to_dev(np.array([3, -4, 5, 7, -3, 2], np.int32))
        iir.inplace(DirectForm1.default("i32", 1, DEV), xy)
        assert to_np(xy).tolist() == want


def test_kat_identity_state_shift():
    """src/iir/biquad.rs:326-338"""
    st = DirectForm1.from_xy([0.0, 1.0], [[2.0, 3.0]], "f32", DEV)
    y0 = Biquad.identity("f32").process(st, to_dev(np.array([4.0], np.float32)))
    assert to_np(y0)[0] == 4.0
    assert st.numpy()[:, 0].tolist() == [4.0, 0.0, 4.0, 2.0]


def test_kat_dither():
    """src/iir/biquad.rs:493-510"""
    st = DirectForm1Dither.from_xye([1, 2], [[3, 4]], 5, DEV)
    y0 = Biquad.identity(Q32(30)).process(st, to_dev(np.array([6], np.int32)))
    assert to_np(y0)[0] == 6
    assert st.numpy()[:, 0].tolist() == [6, 1, 6, 3, 5]


def test_kat_clamp_identity_hold():
    """src/iir/biquad.rs:130-155, 176-212"""
    z = to_dev(np.array([0.0], np.float32))
    i = BiquadClamp(Biquad([0, 0, 0, 0, 0], "f32")); i.u = 5.0
    assert to_np(i.process(DirectForm1.default("f32", 1, DEV), z))[0] == 5.0
    i = BiquadClamp(Biquad([0, 0, 0, 0, 0], "f32")); i.min = 5.0
    assert to_np(i.process(DirectForm1.default("f32", 1, DEV), z))[0] == 5.0
    i = BiquadClamp(Biquad([0, 0, 0, 0, 0], "f32")); i.max = -5.0
    assert to_np(i.process(DirectForm1.default("f32", 1, DEV), z))[0] == -5.0
    x = to_dev(np.array([3.0], np.float32))
    assert to_np(Biquad.identity("f32").process(DirectForm1.default("f32", 1, DEV), x))[0] == 3.0
    assert to_np(Biquad.proportional(2.0, "f32").process(DirectForm1.default("f32", 1, DEV), x))[0] == 6.0
    st = DirectForm1.default("f32", 1, DEV); st.set_y(2.0)
    assert to_np(Biquad.hold("f32").process(st, to_dev(np.array([7.0], np.float32))))[0] == 2.0
    b = BiquadClamp(Biquad.identity("f32"))
    assert to_np(b.process(DirectForm2Transposed.default("f32", 1, DEV), x))[0] == 3.0


# ------------------------------------------------------------------ DF1 all types / layouts / shapes
SHAPES = [(1, 1), (1, 37), (33, 1), (100, 33), (64, 128), (17, 260), (130, 96), (400, 70), (50, 272)]  # (frames, lanes); 400 frames = several 128-byte tiles + a tail for every sample size


@pytest.mark.parametrize("kind", ["i8", "i16", "i32", "i64", "f32", "f64"])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("clamp", [False, True])
def test_df1_vs_oracle(oracle, kind, layout, clamp):
    rng = np.random.default_rng(sum(map(ord, kind)) * 7 + layout * 3 + int(clamp))
    bq = _coeffs(kind, rng)
    cl = None
    if clamp:
        if kind in BITS:
            lim = 1 << (BITS[kind] - 4)
            cl = [3, -lim, lim]
        else:
            cl = [0.25, -0.5, 0.5]
        bq = BiquadClamp(bq, *cl)
    for frames, lanes in SHAPES:
        x = rand_samples(rng, kind, frames * lanes).reshape(frames, lanes)
        flat = layout_flat(x, layout)
        st0 = rand_samples(rng, kind, 4 * lanes, amp_bits=(BITS[kind] - 4) if kind in BITS else None).reshape(4, lanes)
        so = st0.copy()
        want = oracle.biquad_lanes("df1", kind, bq.ba, bq.F, cl, so, flat, lanes, layout)
        st = DirectForm1(to_dev(st0), kind)
        y = torch.empty_like(to_dev(flat))
        Lanes(bq).block(st, to_dev(flat), y, layout)
        assert_bits_equal(to_np(y), want, f"{kind} layout={layout} {frames}x{lanes}")
        assert_bits_equal(st.numpy(), so, "state")


@pytest.mark.parametrize("kind", ["i8", "i16", "i32", "i64"])
def test_df1_fractional_bits_sweep(oracle, kind):
    """every quantiser variant: F = 0, the usual range, the top of it, beyond the sample width and negative
    (generic-shift kernels), with full-scale raw coefficients and samples so the wide accumulator wraps
    and every carry of the i64 (i128 accumulator) chain is exercised"""
    bits = BITS[kind]
    rng = np.random.default_rng(bits)
    frames, lanes = 48, 64
    info = np.iinfo(NP[kind])
    for F in (0, 1, bits // 2, bits - 2, bits - 1, bits, bits + 3, 2 * bits - 1, -1, -(bits - 1)):
        for clamp in (None, [3, int(info.min) // 3, int(info.max) // 5]):
            ba = rng.integers(info.min, info.max, 5, dtype=NP[kind], endpoint=True)
            x = rng.integers(info.min, info.max, frames * lanes, dtype=NP[kind], endpoint=True)
            x[:8] = [info.min, info.max, -1, 0, 1, info.min, info.min, info.max]
            st0 = rng.integers(info.min, info.max, (4, lanes), dtype=NP[kind], endpoint=True)
            for layout in (0, 1):
                so = st0.copy()
                want = oracle.biquad_lanes("df1", kind, ba, F, clamp, so, x, lanes, layout)
                st = DirectForm1(to_dev(st0), kind)
                y = torch.empty_like(to_dev(x))
                bq = Biquad(ba, Q(kind, F))
                if clamp is not None:
                    bq = BiquadClamp(bq, *clamp)
                Lanes(bq).block(st, to_dev(x), y, layout)
                assert_bits_equal(to_np(y), want, f"{kind} F={F} clamp={clamp is not None} layout={layout}")
                assert_bits_equal(st.numpy(), so, f"state {kind} F={F}")


@pytest.mark.parametrize("kind", ["i32", "f32"])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("policy", [1, 2])
def test_df1_tma_and_generic_kernels(oracle, kind, layout, policy):
    """both kernel families (generic LDG = 1, TMA = 2) give the oracle's bits, incl. ragged tails"""
    rng = np.random.default_rng(policy * 10 + layout)
    bq = _coeffs(kind, rng)
    ctx = ib.default_context(0)
    ctx.set_kernel_policy(policy)
    try:
        for frames, lanes in [(16, 32), (64, 128), (100, 36), (52, 260), (1000, 64), (36, 4)]:
            x = rand_samples(rng, kind, frames * lanes).reshape(frames, lanes)
            flat = layout_flat(x, layout)
            so = np.zeros((4, lanes), NP[kind])
            want = oracle.biquad_lanes("df1", kind, bq.ba, bq.F, None, so, flat, lanes, layout)
            st = DirectForm1.default(kind, lanes, DEV)
            y = torch.empty_like(to_dev(flat))
            Lanes(bq).block(st, to_dev(flat), y, layout)
            assert_bits_equal(to_np(y), want, f"policy={policy} {kind} layout={layout} {frames}x{lanes}")
            assert_bits_equal(st.numpy(), so, "state")
    finally:
        ctx.set_kernel_policy(0)


@pytest.mark.parametrize("layout", [0, 1])
def test_df1_streaming_and_inplace(oracle, layout):
    """state carried across block() calls == one long call; x may alias y (SplitInplace)"""
    rng = np.random.default_rng(5)
    bq = _coeffs("i32", rng)
    lanes, frames = 96, 160
    x = rand_samples(rng, "i32", frames * lanes).reshape(frames, lanes)
    so = np.zeros((4, lanes), np.int32)
    want = oracle.biquad_lanes("df1", "i32", bq.ba, bq.F, None, so, layout_flat(x, layout), lanes, layout)
    want = want.reshape(frames, lanes) if layout == 0 else want.reshape(lanes, frames).T
    st = DirectForm1.default("i32", lanes, DEV)
    outs = []
    for a, b in ((0, 48), (48, 49), (49, 160)):
        xy = to_dev(layout_flat(x[a:b], layout))
        Lanes(bq).inplace(st, xy, layout)
        o = to_np(xy)
        outs.append(o.reshape(b - a, lanes) if layout == 0 else o.reshape(lanes, b - a).T)
    assert_bits_equal(np.concatenate(outs), want)
    assert_bits_equal(st.numpy(), so)


def test_process_view_lane_major(oracle):
    """Lanes::process_view on View<LaneMajor> (compose.rs:478-494)"""
    rng = np.random.default_rng(6)
    bq = _coeffs("i32", rng)
    lanes, frames = 5, 64
    x = rand_samples(rng, "i32", lanes * frames)
    p = Split.new(bq, DirectForm1.default("i32", 1, DEV)).lanes(lanes)
    y = torch.empty(lanes * frames, dtype=torch.int32, device=DEV)
    p.process_view(View.from_flat(to_dev(x), frames, lanes), View.from_flat(y, frames, lanes))
    so = np.zeros((4, lanes), np.int32)
    assert_bits_equal(to_np(y), oracle.biquad_lanes("df1", "i32", bq.ba, bq.F, None, so, x, lanes, 1))


def test_df1_host_buffers(oracle):
    """numpy in/out goes through idsp_biquad_df1_*_host (H2D/D2H inside the C ABI)"""
    rng = np.random.default_rng(8)
    bq = _coeffs("i32", rng)
    for layout in (0, 1):
        lanes, frames = 260, 300
        x = rand_samples(rng, "i32", frames * lanes)
        so = np.zeros((4, lanes), np.int32)
        want = oracle.biquad_lanes("df1", "i32", bq.ba, bq.F, None, so, x, lanes, layout)
        st = DirectForm1.default("i32", lanes, None)
        y = np.empty_like(x)
        Lanes(bq).block(st, x, y, layout)
        assert_bits_equal(y, want)
        assert_bits_equal(st.words, so)


# ------------------------------------------------------------------ other forms
@pytest.mark.parametrize("kind", ["f32", "f64"])
@pytest.mark.parametrize("layout", [0, 1])
def test_df2t_vs_oracle(oracle, kind, layout):
    rng = np.random.default_rng(12)
    bq = _coeffs(kind, rng)
    for clamp in (None, [0.1, -0.7, 0.7]):
        cfg = bq if clamp is None else BiquadClamp(bq, *clamp)
        for frames, lanes in [(50, 70), (1, 3), (33, 128)]:
            x = rand_samples(rng, kind, frames * lanes)
            so = np.zeros((2, lanes), NP[kind])
            want = oracle.biquad_lanes("df2t", kind, bq.ba, 0, clamp, so, x, lanes, layout)
            st = DirectForm2Transposed.default(kind, lanes, DEV)
            y = torch.empty_like(to_dev(x))
            Lanes(cfg).block(st, to_dev(x), y, layout)
            assert_bits_equal(to_np(y), want)
            assert_bits_equal(st.numpy(), so)


@pytest.mark.parametrize("form,State", [("df1wide", DirectForm1Wide), ("df1dither", DirectForm1Dither)])
@pytest.mark.parametrize("layout", [0, 1])
def test_wide_dither_vs_oracle(oracle, form, State, layout):
    rng = np.random.default_rng(13)
    bq = Biquad.from_ba6(Filter().critical_frequency(0.02).lowpass(), Q32(29))
    for clamp in (None, [17, -(1 << 28), 1 << 28]):
        cfg = bq if clamp is None else BiquadClamp(bq, *clamp)
        for frames, lanes in [(200, 40), (64, 64), (3, 1)]:
            x = rand_samples(rng, "i32", frames * lanes, 30)
            W = State.default(1).words.shape[0]
            so = np.zeros((W, lanes), np.int32)
            want = oracle.biquad_lanes(form, "i32", bq.ba, 29, clamp, so, x, lanes, layout)
            st = State.default(lanes, DEV)
            y = torch.empty_like(to_dev(x))
            Lanes(cfg).block(st, to_dev(x), y, layout)
            assert_bits_equal(to_np(y), want, form)
            assert_bits_equal(st.numpy(), so, form + " state")


@pytest.mark.parametrize("kind", ["i32", "f32", "i16", "f64"])
@pytest.mark.parametrize("nsec", [1, 3, 4, 7])
def test_cascade_vs_oracle(oracle, kind, nsec):
    rng = np.random.default_rng(14 + nsec)
    secs = []
    for i in range(nsec):
        ba6 = Filter().critical_frequency(0.03 + 0.02 * i).lowpass()
        secs.append(Biquad.from_ba6(ba6, Q(kind, BITS[kind] - 3) if kind in BITS else kind))
    ba = np.stack([s.ba for s in secs])
    F = secs[0].F
    for layout, (frames, lanes) in [(l, s) for l in (0, 1) for s in ((77, 45), (76, 40))]:  # 76x40: TMA tiles in both layouts
        x = rand_samples(rng, kind, frames * lanes, (BITS[kind] - 6) if kind in BITS else None)
        so = np.zeros((2 + 2 * nsec, lanes), NP[kind])
        want = oracle.biquad_lanes("cascade", kind, ba, F, None, so, x, lanes, layout, nsec=nsec)
        st = DirectForm.default(nsec, kind, lanes, DEV)
        y = torch.empty_like(to_dev(x))
        Lanes(Cascade(secs)).block(st, to_dev(x), y, layout)
        assert_bits_equal(to_np(y), want)
        assert_bits_equal(st.numpy(), so)


def test_python_ffi_sos(oracle):
    """idsp.sos / idsp.sos_clamp_wide (src/py.rs:50-108) on device and host arrays"""
    rng = np.random.default_rng(15)
    rows = np.array([sum(Filter().critical_frequency(f).lowpass(), []) for f in (0.05, 0.1, 0.2)])
    x = rand_samples(rng, "i32", 500, 24)
    want = x.copy(); oracle.sos(rows, want)
    xy = to_dev(x); ib.sos(rows, xy)
    assert_bits_equal(to_np(xy), want)
    xh = x.copy(); ib.sos(rows, xh)
    assert_bits_equal(xh, want)
    rows9 = np.concatenate([rows, np.array([[3.4, -1e6, 1e6], [-2.5, -5e5, 5e5], [0.5, -1e5, 2e5]])], axis=1)
    want = x.copy(); oracle.sos_clamp_wide(rows9, want)
    xy = to_dev(x); ib.sos_clamp_wide(rows9, xy)
    assert_bits_equal(to_np(xy), want)
    with pytest.raises(TypeError):
        ib.sos(rows9, to_dev(x))


def test_bad_arguments_are_errors():
    from idsp_b200._lib import IdspError
    st = DirectForm1.default("i32", 4, DEV)
    x = to_dev(np.zeros(16, np.int32))
    with pytest.raises(IdspError):
        Biquad([1, 0, 0, 0, 0], Q32(64)).block(st, x, x)  # F out of range
    with pytest.raises(IdspError):
        Biquad([1, 0, 0, 0, 0], Q32(32)).block(DirectForm1Wide.default(4, DEV), x, x)


@pytest.mark.parametrize("kind", ["i32", "f32"])
@pytest.mark.parametrize("lanes", [9476, 18948, 37892 + 8])
def test_df1_wide_tma_boxes(oracle, kind, lanes):
    """large lane counts select the wide-box TMA configurations (64 / 128 / 256 lanes per CTA);
    ragged lane and frame tails are zero-filled / clipped by the TMA unit"""
    rng = np.random.default_rng(lanes)
    bq = _coeffs(kind, rng)
    frames = 27
    x = rand_samples(rng, kind, frames * lanes)
    st0 = rand_samples(rng, kind, 4 * lanes, 20 if kind in BITS else None).reshape(4, lanes)
    so = st0.copy()
    want = oracle.biquad_lanes("df1", kind, bq.ba, bq.F, None, so, x, lanes, 0, nthreads=4)
    st = DirectForm1(to_dev(st0), kind)
    y = torch.empty_like(to_dev(x))
    Lanes(bq).block(st, to_dev(x), y, 0)
    assert_bits_equal(to_np(y), want)
    assert_bits_equal(st.numpy(), so)


# ------------------------------------------------------------------ 8-byte samples on the TMA kernels (round 2)
@pytest.mark.parametrize("kind", ["i64", "f64"])
@pytest.mark.parametrize("layout", [0, 1])
def test_df1_8byte_tma_and_generic(oracle, kind, layout):
    """i64 / f64 DF1 through the generic kernels (policy 1) and the TMA kernels (policy 2: 8-byte boxes,
    128-byte swizzled lane-major rows), incl. ragged boxes and partial last tiles"""
    rng = np.random.default_rng(layout + 11)
    bq = _coeffs(kind, rng)
    ctx = ib.default_context(0)
    for policy in (1, 2):
        ctx.set_kernel_policy(policy)
        try:
            for frames, lanes in [(16, 32), (64, 128), (100, 36), (52, 260), (1000, 64), (36, 4), (24, 1024)]:
                x = rand_samples(rng, kind, frames * lanes).reshape(frames, lanes)
                flat = layout_flat(x, layout)
                st0 = rand_samples(rng, kind, 4 * lanes, amp_bits=(BITS[kind] - 4) if kind in BITS else None).reshape(4, lanes)
                so = st0.copy()
                want = oracle.biquad_lanes("df1", kind, bq.ba, bq.F, None, so, flat, lanes, layout)
                st = DirectForm1(to_dev(st0), kind)
                y = torch.empty_like(to_dev(flat))
                Lanes(bq).block(st, to_dev(flat), y, layout)
                assert_bits_equal(to_np(y), want, f"policy={policy} {kind} layout={layout} {frames}x{lanes} [{ctx.last_kernel}]")
                assert_bits_equal(st.numpy(), so, "state")
                if policy == 2:
                    assert ctx.last_kernel.startswith("tma"), ctx.last_kernel
        finally:
            ctx.set_kernel_policy(0)


def test_last_kernel_reports_the_fallback():
    """the silent fall-back from the TMA kernels (misaligned view) is observable through the ABI"""
    ctx = ib.default_context(0)
    bq = _coeffs("i32", None)
    lanes, frames = 64, 64
    x = torch.zeros(frames * lanes + 1, dtype=torch.int32, device=DEV)
    y = torch.zeros(frames * lanes + 1, dtype=torch.int32, device=DEV)
    Lanes(bq).block(DirectForm1.default("i32", lanes, DEV), x[:-1], y[:-1], 0)
    assert ctx.last_kernel.startswith("tma frame-major"), ctx.last_kernel
    Lanes(bq).block(DirectForm1.default("i32", lanes, DEV), x[1:], y[1:], 0)
    assert ctx.last_kernel == "generic frame-major", ctx.last_kernel


@pytest.mark.parametrize("clamp", [False, True])
def test_df1_i8_packed_four_lanes_per_word(oracle, monkeypatch, clamp):
    """i8 frame-major from 2^18 lanes on (here forced by IDSP_I8_PACKED_MIN_LANES, read per call): four adjacent
    lanes share a 32-bit word and ride the tensor-map kernels (PackedOp<Df1Op<i8>>); state stays one word per
    lane.  Same bits as the oracle and as the generic kernel."""
    import os

    rng = np.random.default_rng(88 + clamp)
    kind, lanes, frames = "i8", 4096 + 16, 37
    bq = _coeffs(kind, rng)
    cl = [3, -8, 8] if clamp else None
    x = rand_samples(rng, kind, frames * lanes)
    st0 = rand_samples(rng, kind, 4 * lanes, amp_bits=BITS[kind] - 4).reshape(4, lanes)
    so = st0.copy()
    want = oracle.biquad_lanes("df1", kind, bq.ba, bq.F, cl, so, x, lanes, 0, nthreads=4)
    cfg = BiquadClamp(bq, *cl) if clamp else bq
    ctx = ib.default_context(0)
    for force, family in (("1", "tma"), (str(1 << 40), "generic")):
        monkeypatch.setenv("IDSP_I8_PACKED_MIN_LANES", force)
        so = st0.copy()
        oracle.biquad_lanes("df1", kind, bq.ba, bq.F, cl, so, x, lanes, 0, nthreads=4)
        st = DirectForm1(to_dev(st0), kind)
        y = torch.empty_like(to_dev(x))
        Lanes(cfg).block(st, to_dev(x), y, 0)
        assert ctx.last_kernel.startswith(family), ctx.last_kernel
        assert_bits_equal(to_np(y), want, f"packed forced={force}")
        assert_bits_equal(st.numpy(), so)
