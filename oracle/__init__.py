"""ctypes/numpy wrapper around the CPU oracle (oracle/idsp_oracle.c).

TEST INFRASTRUCTURE ONLY.  Nothing under ``idsp_b200/`` may import this module.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, as the checker or the timed CPU baseline.

The oracle restates the reference's arithmetic in C; see idsp_oracle.h for the
reference file:line citations of every function.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libidsp_oracle.so")

FRAME_MAJOR = 0
LANE_MAJOR = 1

_NP = {
    "i8": np.int8,
    "i16": np.int16,
    "i32": np.int32,
    "i64": np.int64,
    "f32": np.float32,
    "f64": np.float64,
}
_CT = {
    "i8": C.c_int8,
    "i16": C.c_int16,
    "i32": C.c_int32,
    "i64": C.c_int64,
    "f32": C.c_float,
    "f64": C.c_double,
}


def build_native() -> str:
    """Rebuild the oracle for the host it runs on (`-march=native`, SURVEY 8(d) "CPU side-by-side") into
    oracle/_native/ and make it the library this process uses; used by bench.py's CPU legs on the box they
    are timed on.  Falls back to the portable build (x86-64-v3) if the compiler is missing."""
    global _SO, _lib
    out_dir = os.path.join(_HERE, "_native")
    so = os.path.join(out_dir, "libidsp_oracle.so")
    try:
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC",
                               "-std=gnu11", "-shared", "-o", so, os.path.join(_HERE, "idsp_oracle.c"), "-lm"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        return build()
    # switch this process to the native build even if the portable one is already loaded (a second dlopen of a
    # different file: the C library is stateless, so calls made before and after give the same results)
    if _SO != so:
        _SO, _lib = so, None
    return so


def build_flags() -> str:
    return "-O3 -march=native" if _SO.endswith(os.path.join("_native", "libidsp_oracle.so")) else "-O3 -march=x86-64-v3"


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("idsp_oracle.c", "idsp_oracle.h", "Makefile")]
    if (
        force
        or not os.path.exists(_SO)
        or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    ):
        subprocess.check_call(["make", "-C", _HERE, "CC=gcc", "-s"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_hbf_dec_state_words.restype = C.c_size_t
        _lib.orc_hbf_int_state_words.restype = C.c_size_t
        _lib.orc_hbf_dec_response_length.restype = C.c_size_t
        _lib.orc_hbf_int_response_length.restype = C.c_size_t
        _lib.orc_cossin_table.restype = C.POINTER(C.c_uint32)
        _lib.orc_atan2_divi_base.restype = C.POINTER(C.c_uint32)
        _lib.orc_atan2_divi_slope.restype = C.POINTER(C.c_int32)
        _lib.orc_atan2.restype = C.c_int32
        for s in ("i8", "i16", "i32", "i64"):
            getattr(_lib, f"orc_quant_{s}").restype = _CT[s]
            getattr(_lib, f"orc_quant_{s}").argtypes = [C.c_double, C.c_int]
        _lib.orc_round_sat_i32.restype = C.c_int32
        _lib.orc_round_sat_i32.argtypes = [C.c_double]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dt, copy=False):
    a = np.asarray(a, dtype=dt)
    if copy or not a.flags["C_CONTIGUOUS"]:
        a = np.array(a, dtype=dt, order="C", copy=True)
    return a


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---------------------------------------------------------------- tables
def cossin_table() -> np.ndarray:
    return np.ctypeslib.as_array(lib().orc_cossin_table(), shape=(128,)).copy()


def atan2_divi_table():
    b = np.ctypeslib.as_array(lib().orc_atan2_divi_base(), shape=(16,)).copy()
    s = np.ctypeslib.as_array(lib().orc_atan2_divi_slope(), shape=(16,)).copy()
    return b, s


# ---------------------------------------------------------------- memoryless
def cossin(phase) -> np.ndarray:
    p = _arr(phase, np.int32).ravel()
    out = np.empty((p.size, 2), np.int32)
    lib().orc_cossin_n(_p(p), _p(out), C.c_size_t(p.size))
    return out


def atan2(xy) -> np.ndarray:
    """xy: [n,2] = (x, y) rows like idsp.atan2 (src/py.rs:31-46)."""
    xy = _arr(xy, np.int32).reshape(-1, 2)
    out = np.empty(xy.shape[0], np.int32)
    lib().orc_atan2_n(_p(xy), _p(out), C.c_size_t(xy.shape[0]))
    return out


# ---------------------------------------------------------------- coefficients
def quantize(v, F: int, kind: str = "i32"):
    f = getattr(lib(), f"orc_quant_{kind}")
    return np.array([f(float(x), int(F)) for x in np.atleast_1d(v)], dtype=_NP[kind])


def ba_normalize(ba6) -> np.ndarray:
    ba6 = _arr(np.asarray(ba6, np.float64).ravel(), np.float64)
    out = np.empty(5, np.float64)
    lib().orc_ba_normalize_f64(_p(ba6), _p(out))
    return out


def ba_from_sos_row(row6, F: int, kind: str = "i32") -> np.ndarray:
    """[[b0,b1,b2],[a0,a1,a2]] -> Biquad<Q<..,F>> raw coefficients (biquad.rs:545-576)."""
    n5 = ba_normalize(row6)
    if kind in ("f32", "f64"):
        return n5.astype(_NP[kind])
    return quantize(n5, F, kind)


def round_sat_i32(v) -> int:
    return int(lib().orc_round_sat_i32(float(v)))


def _clamp(cl, kind):
    return None if cl is None else _arr(cl, _NP[kind], copy=True)


# ---------------------------------------------------------------- biquads (single lane)
def biquad_df1(kind, ba, F, clamp, st, x):
    """st: np array [4] = [x0,x1,y0,y1], updated in place. Returns y."""
    dt = _NP[kind]
    ba = _arr(ba, dt, copy=True)
    x = _arr(x, dt)
    y = np.empty_like(x)
    assert st.dtype == dt and st.size == 4
    cl = _clamp(clamp, kind)
    getattr(lib(), f"orc_biquad_df1_{kind}")(_p(ba), C.c_int(F), _p(cl), _p(st), _p(x), _p(y), C.c_size_t(x.size))
    return y


def biquad_df2t(kind, ba, clamp, st, x):
    dt = _NP[kind]
    ba = _arr(ba, dt, copy=True)
    x = _arr(x, dt)
    y = np.empty_like(x)
    assert st.dtype == dt and st.size == 2
    cl = _clamp(clamp, kind)
    getattr(lib(), f"orc_biquad_df2t_{kind}")(_p(ba), _p(cl), _p(st), _p(x), _p(y), C.c_size_t(x.size))
    return y


def biquad_cascade(kind, ba, F, st, x):
    dt = _NP[kind]
    ba = _arr(ba, dt, copy=True).reshape(-1, 5)
    x = _arr(x, dt)
    y = np.empty_like(x)
    assert st.dtype == dt and st.size == 2 + 2 * ba.shape[0]
    getattr(lib(), f"orc_biquad_cascade_{kind}")(_p(ba), C.c_int(F), C.c_int(ba.shape[0]), _p(st), _p(x), _p(y), C.c_size_t(x.size))
    return y


def biquad_df1wide(ba, F, clamp, st, x):
    ba = _arr(ba, np.int32, copy=True)
    x = _arr(x, np.int32)
    y = np.empty_like(x)
    assert st.dtype == np.int32 and st.size == 6
    cl = _clamp(clamp, "i32")
    lib().orc_biquad_df1wide_i32(_p(ba), C.c_int(F), _p(cl), _p(st), _p(x), _p(y), C.c_size_t(x.size))
    return y


def biquad_df1dither(ba, F, clamp, st, x):
    ba = _arr(ba, np.int32, copy=True)
    x = _arr(x, np.int32)
    y = np.empty_like(x)
    assert st.dtype == np.int32 and st.size == 5
    cl = _clamp(clamp, "i32")
    lib().orc_biquad_df1dither_i32(_p(ba), C.c_int(F), _p(cl), _p(st), _p(x), _p(y), C.c_size_t(x.size))
    return y


# ---------------------------------------------------------------- biquads (lanes)
def _fl(x, layout, lanes):
    """(frames, lanes) of a flat or 2-D array in the given layout."""
    n = x.size
    assert n % lanes == 0
    return n // lanes


def biquad_lanes(form, kind, ba, F, clamp, st, x, lanes, layout=FRAME_MAJOR, nthreads=1, nsec=1):
    """form in {df1, df2t, df1wide, df1dither, cascade}; st: [words, lanes] SoA, in place."""
    dt = _NP[kind]
    ba = _arr(ba, dt, copy=True)
    x = _arr(x, dt)
    y = np.empty_like(x)
    frames = _fl(x, layout, lanes)
    cl = _clamp(clamp, kind)
    L = lib()
    sz = (C.c_size_t(frames), C.c_size_t(lanes), C.c_int(layout), C.c_int(nthreads))
    assert st.flags["C_CONTIGUOUS"]
    if form == "df1":
        assert st.dtype == dt and st.shape == (4, lanes)
        getattr(L, f"orc_biquad_df1_{kind}_lanes")(_p(ba), C.c_int(F), _p(cl), _p(st), _p(x), _p(y), *sz)
    elif form == "df2t":
        assert st.dtype == dt and st.shape == (2, lanes)
        getattr(L, f"orc_biquad_df2t_{kind}_lanes")(_p(ba), _p(cl), _p(st), _p(x), _p(y), *sz)
    elif form == "df1wide":
        assert kind == "i32" and st.dtype == np.int32 and st.shape == (6, lanes)
        L.orc_biquad_df1wide_i32_lanes(_p(ba), C.c_int(F), _p(cl), _p(st), _p(x), _p(y), *sz)
    elif form == "df1dither":
        assert kind == "i32" and st.dtype == np.int32 and st.shape == (5, lanes)
        L.orc_biquad_df1dither_i32_lanes(_p(ba), C.c_int(F), _p(cl), _p(st), _p(x), _p(y), *sz)
    elif form == "cascade":
        assert st.dtype == dt and st.shape == (2 + 2 * nsec, lanes)
        getattr(L, f"orc_biquad_cascade_{kind}_lanes")(_p(ba), C.c_int(F), C.c_int(nsec), _p(st), _p(x), _p(y), *sz)
    else:
        raise ValueError(form)
    return y


# ---------------------------------------------------------------- python FFI shaped
def sos(sos_rows, xy):
    s = _arr(sos_rows, np.float64).reshape(-1, 6)
    assert xy.dtype == np.int32 and xy.flags["C_CONTIGUOUS"]
    lib().orc_sos(_p(s), C.c_int(s.shape[0]), _p(xy), C.c_size_t(xy.size))


def sos_clamp_wide(sos_rows, xy):
    s = _arr(sos_rows, np.float64).reshape(-1, 9)
    assert xy.dtype == np.int32 and xy.flags["C_CONTIGUOUS"]
    lib().orc_sos_clamp_wide(_p(s), C.c_int(s.shape[0]), _p(xy), C.c_size_t(xy.size))


# ---------------------------------------------------------------- hbf
def hbf_taps(idx: int) -> np.ndarray:
    m = C.c_int(0)
    lib().orc_hbf_taps.restype = C.POINTER(C.c_float)
    p = lib().orc_hbf_taps(C.c_int(idx), C.byref(m))
    return np.ctypeslib.as_array(p, shape=(m.value,)).copy()


def hbf_dec_state_words(k):
    return int(lib().orc_hbf_dec_state_words(C.c_int(k)))


def hbf_int_state_words(k):
    return int(lib().orc_hbf_int_state_words(C.c_int(k)))


def hbf_dec_response_length(k):
    return int(lib().orc_hbf_dec_response_length(C.c_int(k)))


def hbf_int_response_length(k):
    return int(lib().orc_hbf_int_response_length(C.c_int(k)))


def hbf_dec(taps, st, x):
    taps = _arr(taps, np.float32, copy=True)
    M = taps.size
    x = _arr(x, np.float32).ravel()
    assert x.size % 2 == 0 and st.dtype == np.float32 and st.size == 3 * M - 2
    y = np.empty(x.size // 2, np.float32)
    lib().orc_hbf_dec_f32(_p(taps), C.c_int(M), _p(st), _p(x), _p(y), C.c_size_t(y.size))
    return y


def hbf_int(taps, st, x):
    taps = _arr(taps, np.float32, copy=True)
    M = taps.size
    x = _arr(x, np.float32).ravel()
    assert st.dtype == np.float32 and st.size == 2 * M - 1
    y = np.empty(x.size * 2, np.float32)
    lib().orc_hbf_int_f32(_p(taps), C.c_int(M), _p(st), _p(x), _p(y), C.c_size_t(x.size))
    return y


def fir(taps, odd, sym, st, x):
    taps = _arr(taps, np.float32, copy=True)
    M = taps.size
    x = _arr(x, np.float32).ravel()
    assert st.dtype == np.float32 and st.size == 2 * M - 1 + int(odd)
    y = np.empty_like(x)
    lib().orc_fir_f32(_p(taps), C.c_int(M), C.c_int(odd), C.c_int(sym), _p(st), _p(x), _p(y), C.c_size_t(x.size))
    return y


def hbf_dec_cascade(k, st, x):
    x = _arr(x, np.float32).ravel()
    R = 1 << k
    assert x.size % R == 0 and st.dtype == np.float32 and st.size == hbf_dec_state_words(k)
    y = np.empty(x.size // R, np.float32)
    lib().orc_hbf_dec_cascade_f32(C.c_int(k), _p(st), _p(x), _p(y), C.c_size_t(y.size))
    return y


def hbf_int_cascade(k, st, x):
    x = _arr(x, np.float32).ravel()
    R = 1 << k
    assert st.dtype == np.float32 and st.size == hbf_int_state_words(k)
    y = np.empty(x.size * R, np.float32)
    lib().orc_hbf_int_cascade_f32(C.c_int(k), _p(st), _p(x), _p(y), C.c_size_t(x.size))
    return y


def hbf_dec_cascade_lanes(k, st, x, lanes, layout=FRAME_MAJOR, nthreads=1):
    x = _arr(x, np.float32).ravel()
    R = 1 << k
    n_out = x.size // (R * lanes)
    assert st.shape == (hbf_dec_state_words(k), lanes) and st.dtype == np.float32
    y = np.empty(n_out * lanes, np.float32)
    lib().orc_hbf_dec_cascade_f32_lanes(C.c_int(k), _p(st), _p(x), _p(y), C.c_size_t(n_out), C.c_size_t(lanes), C.c_int(layout), C.c_int(nthreads))
    return y


def hbf_int_cascade_lanes(k, st, x, lanes, layout=FRAME_MAJOR, nthreads=1):
    x = _arr(x, np.float32).ravel()
    R = 1 << k
    n_in = x.size // lanes
    assert st.shape == (hbf_int_state_words(k), lanes) and st.dtype == np.float32
    y = np.empty(n_in * lanes * R, np.float32)
    lib().orc_hbf_int_cascade_f32_lanes(C.c_int(k), _p(st), _p(x), _p(y), C.c_size_t(n_in), C.c_size_t(lanes), C.c_int(layout), C.c_int(nthreads))
    return y


# ---------------------------------------------------------------- lowpass / lockin
def lowpass(k, st, x):
    k = _arr(k, np.int32, copy=True)
    x = _arr(x, np.int32).ravel()
    assert st.dtype == np.int64 and st.size == k.size
    y = np.empty_like(x)
    lib().orc_lowpass_i32(C.c_int(k.size), _p(k), _p(st), _p(x), _p(y), C.c_size_t(x.size))
    return y


def lowpass_lanes(k, st, x, lanes, layout=FRAME_MAJOR, nthreads=1):
    k = _arr(k, np.int32, copy=True)
    x = _arr(x, np.int32).ravel()
    frames = x.size // lanes
    assert st.dtype == np.int64 and st.shape == (k.size, lanes)
    y = np.empty_like(x)
    lib().orc_lowpass_i32_lanes(C.c_int(k.size), _p(k), _p(st), _p(x), _p(y), C.c_size_t(frames), C.c_size_t(lanes), C.c_int(layout), C.c_int(nthreads))
    return y


def lockin_lanes(k, accu_state, accu_step, lp_st, x, lanes, layout=FRAME_MAJOR, nthreads=1):
    k = _arr(k, np.int32, copy=True)
    x = _arr(x, np.int32).ravel()
    frames = x.size // lanes
    order = k.size
    assert accu_state.dtype == np.int32 and accu_state.shape == (lanes,)
    accu_step = _arr(accu_step, np.int32)
    assert lp_st.dtype == np.int64 and lp_st.shape == (2 * order, lanes)
    iq = np.empty(x.size * 2, np.int32)
    lib().orc_lockin_i32_lanes(C.c_int(order), _p(k), _p(accu_state), _p(accu_step), _p(lp_st), _p(x), _p(iq), C.c_size_t(frames), C.c_size_t(lanes), C.c_int(layout), C.c_int(nthreads))
    return iq


def lockin_phase_lanes(k, lp_st, xp, lanes, layout=FRAME_MAJOR, nthreads=1):
    """(sample, phase) tuples, src/lockin.rs:30-39; xp = (x, phase) pairs"""
    k = _arr(k, np.int32, copy=True)
    xp = _arr(xp, np.int32).ravel()
    frames = xp.size // (2 * lanes)
    order = k.size
    assert lp_st.dtype == np.int64 and lp_st.shape == (2 * order, lanes)
    iq = np.empty(frames * lanes * 2, np.int32)
    lib().orc_lockin_phase_i32_lanes(C.c_int(order), _p(k), _p(lp_st), _p(xp), _p(iq), C.c_size_t(frames), C.c_size_t(lanes), C.c_int(layout), C.c_int(nthreads))
    return iq


def lockin_lo_lanes(k, lp_st, xlo, lanes, layout=FRAME_MAJOR, nthreads=1):
    """(sample, LO) tuples, src/lockin.rs:17-28; xlo = (x, lo.re, lo.im) triples"""
    k = _arr(k, np.int32, copy=True)
    xlo = _arr(xlo, np.int32).ravel()
    frames = xlo.size // (3 * lanes)
    order = k.size
    assert lp_st.dtype == np.int64 and lp_st.shape == (2 * order, lanes)
    iq = np.empty(frames * lanes * 2, np.int32)
    lib().orc_lockin_lo_i32_lanes(C.c_int(order), _p(k), _p(lp_st), _p(xlo), _p(iq), C.c_size_t(frames), C.c_size_t(lanes), C.c_int(layout), C.c_int(nthreads))
    return iq


def chain_lanes(k, ba, st, x, lanes, layout=FRAME_MAJOR, nthreads=1):
    ba = _arr(ba, np.float32, copy=True)
    x = _arr(x, np.float32).ravel()
    R = 1 << k
    n_low = x.size // (R * lanes)
    W = hbf_dec_state_words(k) + hbf_int_state_words(k) + 4
    assert st.dtype == np.float32 and st.shape == (W, lanes)
    y = np.empty_like(x)
    lib().orc_chain_f32_lanes(C.c_int(k), _p(ba), _p(st), _p(x), _p(y), C.c_size_t(n_low), C.c_size_t(lanes), C.c_int(layout), C.c_int(nthreads))
    return y


# ---------------------------------------------------------------- Cic<T,N,M> (src/cic.rs)
def cic_state_words(N, M):
    return 2 + N * M + N


def _cic(kind_fn, N, M, rate, st, x, lanes, layout, nthreads, R_in, R_out):
    dt = st.dtype
    assert dt in (np.int32, np.int64) and st.shape == (cic_state_words(N, M), lanes)
    x = _arr(x, dt)
    frames = x.size // (lanes * R_in)
    y = np.empty(frames * lanes * R_out, dt)
    fn = getattr(lib(), f"orc_cic_{kind_fn}_{'i32' if dt == np.int32 else 'i64'}_lanes")
    fn(C.c_int(N), C.c_int(M), C.c_uint32(rate), _p(st), _p(x), _p(y), C.c_size_t(frames), C.c_size_t(lanes),
       C.c_int(layout), C.c_int(nthreads))
    return y


def cic_dec_lanes(N, M, rate, st, x, lanes, layout=FRAME_MAJOR, nthreads=1):
    """Decimator(Cic::<T,N,M>::new(rate)): frames of rate+1 inputs -> 1 output; `st` updated in place."""
    return _cic("dec", N, M, rate, st, x, lanes, layout, nthreads, rate + 1, 1)


def cic_int_lanes(N, M, rate, st, x, lanes, layout=FRAME_MAJOR, nthreads=1):
    """Interpolator(Cic::<T,N,M>::new(rate)): 1 input -> frames of rate+1 outputs."""
    return _cic("int", N, M, rate, st, x, lanes, layout, nthreads, 1, rate + 1)


def cic_gain(N, M, rate):
    lib().orc_cic_gain.restype = C.c_int64
    return int(lib().orc_cic_gain(C.c_int(N), C.c_int(M), C.c_uint32(rate)))


def cic_gain_log2(N, M, rate):
    lib().orc_cic_gain_log2.restype = C.c_uint32
    return int(lib().orc_cic_gain_log2(C.c_int(N), C.c_int(M), C.c_uint32(rate)))


def cic_response_length(N, rate):
    lib().orc_cic_response_length.restype = C.c_size_t
    return int(lib().orc_cic_response_length(C.c_int(N), C.c_uint32(rate)))


# ---------------------------------------------------------------- PLL (src/pll.rs)
PLL_WORDS = 9


def pll_from_bandwidth(bw, split=4.0):
    ba = np.zeros(3, np.int32)
    lib().orc_pll_from_bandwidth(C.c_float(bw), C.c_float(split), _p(ba))
    return ba


def pll_lanes(ba, st, x, lanes, layout=FRAME_MAJOR, nthreads=1):
    """PLL::process over lanes; st int32 [9, lanes] updated in place; returns y (output phase)."""
    ba = _arr(ba, np.int32)
    x = _arr(x, np.int32)
    assert st.dtype == np.int32 and st.shape == (PLL_WORDS, lanes)
    y = np.empty_like(x)
    lib().orc_pll_i32_lanes(_p(ba), _p(st), _p(x), _p(y), C.c_size_t(x.size // lanes), C.c_size_t(lanes),
                            C.c_int(layout), C.c_int(nthreads))
    return y


# ---------------------------------------------------------------- FM discriminator (examples/fm_disc.rs)
FM_DISC_WORDS = 7


def fm_disc_lanes(carrier, ba, F, st, x, lanes, layout=FRAME_MAJOR, nthreads=1):
    """x: int32 (re, im) pairs, frames*lanes*2 values; st int32 [7, lanes] updated in place."""
    ba = _arr(ba, np.int32)
    x = _arr(x, np.int32)
    assert st.dtype == np.int32 and st.shape == (FM_DISC_WORDS, lanes)
    frames = x.size // (2 * lanes)
    y = np.empty(frames * lanes, np.int32)
    car = int(carrier) - (1 << 32) if int(carrier) >= (1 << 31) else int(carrier)
    lib().orc_fm_disc_i32_lanes(C.c_int32(car), _p(ba), C.c_int(F), _p(st), _p(x), _p(y), C.c_size_t(frames),
                                C.c_size_t(lanes), C.c_int(layout), C.c_int(nthreads))
    return y
