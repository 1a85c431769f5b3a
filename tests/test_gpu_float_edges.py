"""Floating-point edge cases through the f32 / f64 paths, GPU vs CPU oracle, compared as BIT PATTERNS:
NaN, +-Inf, denormals (the library is built -ftz=false), -0.0, FLT_MAX / DBL_MAX overflow.

What the reference guarantees (and what is therefore compared):
  * every finite / infinite / zero result, including the sign of zero and every denormal, bit for bit;
  * NaN *where* the reference produces NaN.  The NaN payload and sign are not part of Rust's guarantees
    (they differ between x86 SSE -- default NaN 0xFFC00000, operand payloads propagated -- and NVIDIA
    GPUs -- canonical 0x7FFFFFFF / 0xFFF8000000000000), so NaNs compare equal to NaNs.
Reference behaviour exercised: `num_traits::clamp` lets NaN pass (src/iir/biquad.rs:400: compares are
false for NaN), `Clamp::MIN/MAX` for floats are -inf / +inf (src/num.rs:33-42), `Iterator::sum::<f32>()`
folds t0 + t1 + ... so an all-(-0.0) half-band window sums to -0.0 (src/hbf.rs:58-66), no FMA, no FTZ."""
import numpy as np
import pytest
import torch

from gpu_common import DEV, NP, to_dev, to_np

pytestmark = pytest.mark.gpu

import idsp_b200 as ib
from idsp_b200 import (Biquad, BiquadClamp, DirectForm1, DirectForm2Transposed, EvenSymmetric, Filter, HbfDec,
                       HbfDecCascade, HbfInt, HbfIntCascade, Lanes, OddSymmetric)
from idsp_b200.hbf import FirState, _dec_state, _int_state


def assert_bits_equal_nan(a, b, what=""):
    """bit-exact, except that any NaN matches any NaN (payload / sign unspecified by the reference)"""
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    ui = {4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN positions differ at {np.nonzero(na != nb)[0][:8]}"
    bad = np.nonzero((a.view(ui) != b.view(ui)) & ~na)[0]
    assert bad.size == 0, f"{what}: {bad.size} mismatches, first at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]} " \
                          f"({a.view(ui)[bad[:5]]} vs {b.view(ui)[bad[:5]]})"


def specials(kind):
    f = np.finfo(NP[kind])
    tiny = f.smallest_subnormal
    return np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, f.max, -f.max, f.tiny, -f.tiny, tiny, -tiny,
                     tiny * 3, f.tiny / 2, -f.tiny / 4, 1.0, -1.0], NP[kind])


def edge_stream(rng, kind, frames, lanes):
    """[frames, lanes]: lane l % 8 selects the flavour so that a NaN / Inf that enters a recurrence only
    poisons its own lane (lanes are independent, dsp-process/src/compose.rs:472-475)"""
    dt = NP[kind]
    f = np.finfo(dt)
    x = rng.standard_normal((frames, lanes)).astype(dt)
    sp = specials(kind)
    for l in range(lanes):
        fl = l % 8
        t = rng.integers(1, max(frames - 1, 2))
        if fl == 0:                       # pure denormal / signed-zero traffic
            x[:, l] = rng.choice(np.array([0.0, -0.0, f.smallest_subnormal, -f.smallest_subnormal,
                                           f.tiny / 2, -f.tiny / 8, f.tiny, -f.tiny], dt), frames)
        elif fl == 1:                     # tiny normals whose products underflow to denormals
            x[:, l] = (rng.standard_normal(frames) * float(f.tiny) * 4).astype(dt)
        elif fl == 2:                     # one NaN mid-stream
            x[t % frames, l] = np.nan
        elif fl == 3:                     # +Inf then -Inf: Inf - Inf -> NaN inside the recurrence
            x[t % frames, l] = np.inf
            x[(t + 2) % frames, l] = -np.inf
        elif fl == 4:                     # overflow: +-MAX samples
            x[:, l] = rng.choice(np.array([f.max, -f.max, f.max / 2, 1.0], dt), frames)
        elif fl == 5:                     # all -0.0
            x[:, l] = -0.0
        elif fl == 6:                     # random sprinkle of every special value
            idx = rng.integers(0, frames, max(frames // 4, 1))
            x[idx, l] = rng.choice(sp, idx.size)
        # fl == 7: ordinary N(0, 1) lane next to the others
    return x


def flat(x_tl, layout):
    return np.ascontiguousarray(x_tl if layout == 0 else np.swapaxes(x_tl, 0, 1)).reshape(-1)


COEFFS = {
    "lowpass": lambda k: np.asarray(Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), k).ba),
    "tiny": lambda k: (np.array([3.0, -2.0, 1.5, 0.5, -0.25]) * float(np.finfo(NP[k]).tiny)).astype(NP[k]),
    "big": lambda k: np.array([1e30 if k == "f32" else 1e300, -0.5, 0.25, 1.5, -0.75], NP[k]),
    "signed_zero": lambda k: np.array([-0.0, 0.0, -0.0, 0.0, -0.0], NP[k]),
}


@pytest.mark.parametrize("kind", ["f32", "f64"])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("form", ["df1", "df2t"])
@pytest.mark.parametrize("clamp", ["none", "default", "finite", "nan_limits"])
def test_biquad_float_edges(oracle, kind, layout, form, clamp):
    rng = np.random.default_rng(sum(map(ord, kind + form + clamp)) * 4 + layout)
    dt = NP[kind]
    nwords = 4 if form == "df1" else 2
    cl = {"none": None, "default": [0.0, -np.inf, np.inf], "finite": [0.25, -0.5, 0.5],
          "nan_limits": [np.float64(np.nan), -1.0, 1.0]}[clamp]
    # shapes: generic kernels (odd sizes), TMA kernels (aligned, >= 16 frames), ragged last box
    for frames, lanes in ((37, 24), (64, 256), (130, 72), (400, 40)):
        x = edge_stream(rng, kind, frames, lanes)
        for cname, mk in COEFFS.items():
            ba = mk(kind)
            st0 = rng.standard_normal((nwords, lanes)).astype(dt)
            st0[:, 5 % lanes] = -0.0
            so = st0.copy()
            xf = flat(x, layout)
            want = oracle.biquad_lanes(form, kind, ba, 0, cl, so, xf, lanes, layout)
            bq = Biquad(ba, kind)
            if cl is not None:
                bq = BiquadClamp(bq, *cl)
            st = (DirectForm1 if form == "df1" else DirectForm2Transposed)(to_dev(st0), kind)
            y = torch.empty_like(to_dev(xf))
            Lanes(bq).block(st, to_dev(xf), y, layout)
            what = f"{form} {kind} layout={layout} clamp={clamp} coeffs={cname} {frames}x{lanes}"
            assert_bits_equal_nan(to_np(y), want, what)
            assert_bits_equal_nan(st.numpy(), so, "state " + what)


def test_clamp_default_limits_are_infinite():
    """`Clamp::MIN/MAX` for f32 / f64 = -inf / +inf (src/num.rs:33-42): an overflowing junction stays
    +inf under the default clamp instead of being pulled back to FLT_MAX"""
    for kind in ("f32", "f64"):
        b = BiquadClamp(Biquad([2.0, 0, 0, 0, 0], kind))
        assert b.min == -np.inf and b.max == np.inf
        x = to_dev(np.array([np.finfo(NP[kind]).max], NP[kind]))
        y = b.process(DirectForm1.default(kind, 1, DEV), x)
        assert np.isposinf(to_np(y)[0])


def test_clamp_nan_passes():
    """num_traits::clamp (src/iir/biquad.rs:400): `NaN < min` and `NaN > max` are false, NaN is returned"""
    b = BiquadClamp(Biquad([1.0, 0, 0, 0, 0], "f32"), 0.0, -1.0, 1.0)
    y = b.process(DirectForm1.default("f32", 1, DEV), to_dev(np.array([np.nan], np.float32)))
    assert np.isnan(to_np(y)[0])
    y = b.process(DirectForm1.default("f32", 1, DEV), to_dev(np.array([np.inf], np.float32)))
    assert to_np(y)[0] == 1.0


def fir_edge_stream(rng, n, lanes):
    """[lanes, n] f32 with specials sprinkled: a FIR forgets them after its window, so one stream can carry all"""
    f = np.finfo(np.float32)
    x = rng.uniform(-1, 1, (lanes, n)).astype(np.float32)
    sp = specials("f32")
    for l in range(lanes):
        fl = l % 6
        if fl == 0:
            x[l] = rng.choice(np.array([0.0, -0.0, f.smallest_subnormal, -f.smallest_subnormal, f.tiny / 2, -f.tiny], np.float32), n)
        elif fl == 1:
            x[l] = -0.0
        elif fl == 2:
            x[l] = (rng.standard_normal(n) * float(f.tiny) * 8).astype(np.float32)
        elif fl == 3:
            idx = rng.integers(0, n, max(n // 16, 1))
            x[l, idx] = rng.choice(sp, idx.size)
        elif fl == 4:
            x[l] = rng.choice(np.array([f.max, -f.max, f.max / 4], np.float32), n)
    return x


@pytest.mark.parametrize("k", [1, 2, 4, 5])
@pytest.mark.parametrize("layout", [0, 1])
def test_hbf_cascade_float_edges(oracle, k, layout):
    """tiled + generic cascade kernels: 48 lanes x 1024+ inputs cover whole tiles and a ragged tail"""
    rng = np.random.default_rng(100 + 10 * k + layout)
    R = 1 << k
    lanes, n_out = 48, (2048 >> k) + 3
    xl = fir_edge_stream(rng, n_out * R, lanes)                       # [lanes][n]
    x_tlr = np.ascontiguousarray(xl.reshape(lanes, n_out, R).swapaxes(0, 1))  # [t][lane][R]
    xf = x_tlr.reshape(-1) if layout == 0 else xl.reshape(-1)
    so = np.zeros((oracle.hbf_dec_state_words(k), lanes), np.float32)
    so[:, 1] = -0.0
    st = _dec_state(k)(lanes, DEV)
    st.words.copy_(to_dev(so))
    want = oracle.hbf_dec_cascade_lanes(k, so, xf, lanes, layout)
    y = torch.empty(n_out * lanes, dtype=torch.float32, device=DEV)
    Lanes(HbfDecCascade(k)).block(st, to_dev(xf), y, layout)
    assert_bits_equal_nan(to_np(y), want, f"HbfDec /{R} layout={layout}")
    assert_bits_equal_nan(st.numpy(), so, "dec state")
    # interpolator on the decimated stream (specials included)
    si = np.zeros((oracle.hbf_int_state_words(k), lanes), np.float32)
    sti = _int_state(k)(lanes, DEV)
    want_i = oracle.hbf_int_cascade_lanes(k, si, want, lanes, layout)
    yi = torch.empty(n_out * lanes * R, dtype=torch.float32, device=DEV)
    Lanes(HbfIntCascade(k)).block(sti, to_dev(want), yi, layout)
    assert_bits_equal_nan(to_np(yi), want_i, f"HbfInt x{R} layout={layout}")
    assert_bits_equal_nan(sti.numpy(), si, "int state")


def test_hbf_all_negative_zero_window(oracle):
    """src/hbf.rs:58-66: the sum is t0 + t1 + ... (fold from -0.0): all-(-0.0) input -> the sign of the
    output follows the tap signs exactly as in the reference, never a blanket +0.0"""
    taps = ib.hbf_taps()[3]
    n, lanes = 64, 4
    x = np.full(2 * n * lanes, -0.0, np.float32)
    so = np.full((3 * taps.size - 2, lanes), -0.0, np.float32)
    want = np.concatenate([oracle.hbf_dec(taps, so[:, l].copy(), x[: 2 * n]) for l in range(lanes)])
    st = HbfDec(to_dev(so), taps.size)
    y = torch.empty(n * lanes, dtype=torch.float32, device=DEV)
    EvenSymmetric(taps).block(st, to_dev(x), y, 1)
    assert_bits_equal_nan(to_np(y), want, "all -0.0 decimator")
    assert np.all(to_np(y) == 0.0)
    # all-positive taps: (-0 + -0) * c = -0, and -0 + -0 + ... = -0: the output must be -0.0, not +0.0
    tp = np.array([0.25, 0.5], np.float32)
    so = np.full((3 * tp.size - 2, 1), -0.0, np.float32)
    want = oracle.hbf_dec(tp, so[:, 0].copy(), x[: 2 * n])
    assert np.all(np.signbit(want))
    y = torch.empty(n, dtype=torch.float32, device=DEV)
    EvenSymmetric(tp).block(HbfDec(to_dev(so), tp.size), to_dev(x[: 2 * n]), y, 1)
    assert_bits_equal_nan(to_np(y), want, "-0.0 sum sign")


@pytest.mark.parametrize("layout", [0, 1])
def test_single_stage_and_fir_float_edges(oracle, layout):
    rng = np.random.default_rng(7 + layout)
    lanes, n = 12, 96
    f = np.finfo(np.float32)
    for M in (2, 3, 5, 7, 23):   # templated (registers) and run-time tap kernels
        taps = (rng.standard_normal(M) * 0.2).astype(np.float32)
        taps[0] = f.smallest_subnormal * 5      # denormal tap x sample
        xl = fir_edge_stream(rng, 2 * n, lanes)
        xf = flat(np.ascontiguousarray(xl.reshape(lanes, n, 2).swapaxes(0, 1)), 0) if layout == 0 else xl.reshape(-1)
        so = np.zeros((3 * M - 2, lanes), np.float32)
        want_l = [oracle.hbf_dec(taps, so[:, l].copy(), xl[l]) for l in range(lanes)]
        want = np.stack(want_l, 1).reshape(-1) if layout == 0 else np.concatenate(want_l)
        y = torch.empty(n * lanes, dtype=torch.float32, device=DEV)
        EvenSymmetric(taps).block(HbfDec.default(M, lanes, DEV), to_dev(xf), y, layout)
        assert_bits_equal_nan(to_np(y), want, f"hbf_dec M={M}")
        # x2
        xi = fir_edge_stream(rng, n, lanes)
        xif = flat(np.ascontiguousarray(xi.swapaxes(0, 1)), 0) if layout == 0 else xi.reshape(-1)
        so = np.zeros((2 * M - 1, lanes), np.float32)
        want_l = [oracle.hbf_int(taps, so[:, l].copy(), xi[l]) for l in range(lanes)]
        want = np.stack([w.reshape(n, 2) for w in want_l], 1).reshape(-1) if layout == 0 else np.concatenate(want_l)
        y = torch.empty(2 * n * lanes, dtype=torch.float32, device=DEV)
        EvenSymmetric(taps).block(HbfInt.default(M, lanes, DEV), to_dev(xif), y, layout)
        assert_bits_equal_nan(to_np(y), want, f"hbf_int M={M}")
        # single-rate odd-symmetric FIR
        so = np.zeros((2 * M, lanes), np.float32)
        want_l = [oracle.fir(taps, True, True, so[:, l].copy(), xi[l]) for l in range(lanes)]
        want = np.stack(want_l, 1).reshape(-1) if layout == 0 else np.concatenate(want_l)
        y = torch.empty(n * lanes, dtype=torch.float32, device=DEV)
        OddSymmetric(taps).block(FirState.default(2 * M, lanes, DEV), to_dev(xif), y, layout)
        assert_bits_equal_nan(to_np(y), want, f"fir M={M}")


@pytest.mark.parametrize("lanes,n_low", [(16, 64), (8, 256 + 5), (24, 40)])
def test_chain_float_edges(oracle, lanes, n_low):
    """config-5 chain (dec -> int -> biquad) on the fused-biquad tiled path and on the single-pass kernel"""
    from idsp_b200 import _lib

    rng = np.random.default_rng(lanes * 1000 + n_low)
    k, R = 4, 16
    W = int(_lib.lib().idsp_chain_state_words(k))
    ba = np.asarray(Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), "f32").ba)
    xl = fir_edge_stream(rng, n_low * R, lanes)
    ctx = ib.default_context(0)
    for layout in (0, 1):
        xf = np.ascontiguousarray(xl.reshape(lanes, n_low, R).swapaxes(0, 1)).reshape(-1) if layout == 0 else xl.reshape(-1)
        so = np.zeros((W, lanes), np.float32)
        want = oracle.chain_lanes(k, ba, so, xf, lanes, layout)
        st = torch.zeros((W, lanes), dtype=torch.float32, device=DEV)
        y = torch.empty_like(to_dev(xf))
        ctx.chain(k, ba, st, to_dev(xf), y, lanes=lanes, layout=layout)
        assert_bits_equal_nan(to_np(y), want, f"chain layout={layout} {lanes}x{n_low}")
        assert_bits_equal_nan(to_np(st), so, "chain state")
