"""Host mirror of ``idsp::iir`` biquad types on top of the CUDA lane kernels.

``Biquad``/``BiquadClamp``/``Cascade`` are immutable configurations shared by all
lanes; ``DirectForm1`` & co. are per-lane states (SoA device arrays).  Which
kernel runs is selected by the state type, like the reference's trait impls:

  Biquad<Q<T,A,F>> / Biquad<f> on DirectForm1<T>      src/iir/biquad.rs:366-383
  BiquadClamp<C,T>            on DirectForm1<T>       src/iir/biquad.rs:394-404
  Biquad<T> / BiquadClamp<T>  on DirectForm2Transposed src/iir/biquad.rs:418-440
  Biquad<Q32<F>> (+clamp)     on DirectForm1Wide      src/iir/biquad.rs:445-480
  Biquad<Q32<F>> (+clamp)     on DirectForm1Dither    src/iir/biquad.rs:484-538
  Cascade<[Biquad<C>;N]>      on DirectForm<T,N>      src/iir/biquad.rs:339-364
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from .process import LaneState, _Proc

_INT_INFO = {"i8": (np.int8, 8), "i16": (np.int16, 16), "i32": (np.int32, 32), "i64": (np.int64, 64)}
_FLT = {"f32": np.float32, "f64": np.float64}


@dataclass(frozen=True)
class Q:
    """Fixed-point format ``Q<T, A, F>`` (dsp-fixedpoint/src/lib.rs:155-160): base
    type ``kind`` with ``F`` fractional bits; the accumulator is twice as wide."""

    kind: str
    F: int

    @property
    def dtype(self):
        return _INT_INFO[self.kind][0]

    def from_float(self, v: float) -> int:
        """``(v * 2^F).round() as T``: round half away from zero, saturating cast,
        NaN -> 0 (dsp-fixedpoint/src/num_traits_impl.rs:32-45 + Rust `as`)."""
        bits = _INT_INFO[self.kind][1]
        lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
        s = float(v) * math.ldexp(1.0, self.F)
        if s != s:
            return 0
        if math.isinf(s):
            return hi if s > 0 else lo
        a = abs(s)
        r = math.floor(a)
        if a - r >= 0.5:  # exact in binary floating point
            r += 1
        r = int(r) if s >= 0 else -int(r)
        return max(lo, min(hi, r))


def Q8(F): return Q("i8", F)
def Q16(F): return Q("i16", F)
def Q32(F): return Q("i32", F)
def Q64(F): return Q("i64", F)


def _fmt_kind(fmt) -> str:
    return fmt.kind if isinstance(fmt, Q) else fmt


class Biquad(_Proc):
    """``Biquad<C>``: ``ba = [b0, b1, b2, a1, a2]`` (src/iir/biquad.rs:96-116)."""

    def __init__(self, ba: Sequence, fmt):
        """fmt: a :class:`Q` (raw integer coefficients) or 'f32' / 'f64'."""
        self.fmt = fmt
        kind = _fmt_kind(fmt)
        dt = _INT_INFO[kind][0] if kind in _INT_INFO else _FLT[kind]
        self.ba = np.array(ba, dtype=dt).reshape(5)

    # ---- constructors (biquad.rs:545-576)
    @classmethod
    def from_ba6(cls, ba, fmt, src: str = "f64") -> "Biquad":
        """``From<[[f;3];2]>``: literature-sign ``[[b0,b1,b2],[a0,a1,a2]]``.  ``src`` selects the
        reference impl: ``impl_from_float!(f64)`` (default) or ``impl_from_float!(f32)``
        (biquad.rs:545-566), which normalises and quantises in f32 -- e.g. ``Filter<f32>`` ->
        ``Biquad<Q32<30>>`` in examples/fm_disc.rs; it runs through the C ABI builder
        (``idsp_biquad_from_ba6_f32``)."""
        b, a = ba
        kind = _fmt_kind(fmt)
        if src == "f32":
            import ctypes as C

            from . import _lib

            flat = (C.c_float * 6)(*[float(np.float32(v)) for v in (*b, *a)])
            dt = _INT_INFO[kind][0] if kind in _INT_INFO else _FLT[kind]
            out = np.zeros(5, dt)
            _lib.check(_lib.lib().idsp_biquad_from_ba6_f32(flat, _lib.KIND_CODE[kind], fmt.F if isinstance(fmt, Q) else 0,
                                                           out.ctypes.data_as(C.c_void_p)))
            return cls(out, fmt)
        if src != "f64":
            raise ValueError("src must be 'f64' or 'f32'")
        if kind == "f32":  # the f32 impl does the normalisation in f32 (impl_from_float!(f32))
            f = np.float32
            a0 = f(1.0) / f(a[0])
            n5 = [f(b[0]) * a0, f(b[1]) * a0, f(b[2]) * a0, -f(a[1]) * a0, -f(a[2]) * a0]
            return cls(n5, fmt)
        a0 = 1.0 / float(a[0])
        n5 = [b[0] * a0, b[1] * a0, b[2] * a0, -a[1] * a0, -a[2] * a0]
        return cls.from_normalized(n5, fmt)

    @classmethod
    def from_normalized(cls, ba5, fmt) -> "Biquad":
        """``From<[T;5]>``: normalised, sign-flipped ``[b0,b1,b2,a1,a2]`` floats."""
        if isinstance(fmt, Q):
            return cls([fmt.from_float(v) for v in ba5], fmt)
        return cls(ba5, fmt)

    @classmethod
    def proportional(cls, k, fmt) -> "Biquad":
        return cls([k, 0, 0, 0, 0], fmt)

    @classmethod
    def identity(cls, fmt) -> "Biquad":
        """``Biquad::IDENTITY`` (biquad.rs:183)."""
        one = (1 << fmt.F) if isinstance(fmt, Q) else 1.0
        return cls.proportional(one, fmt)

    @classmethod
    def hold(cls, fmt) -> "Biquad":
        """``Biquad::HOLD`` (biquad.rs:210-212)."""
        one = (1 << fmt.F) if isinstance(fmt, Q) else 1.0
        return cls([0, 0, 0, one, 0], fmt)

    def forward_gain(self):
        return self.ba[0] + self.ba[1] + self.ba[2]

    @property
    def F(self) -> int:
        return self.fmt.F if isinstance(self.fmt, Q) else 0

    @property
    def kind(self) -> str:
        return _fmt_kind(self.fmt)

    def _clamp(self):
        return None

    def _block(self, ctx, state, x, y, layout):
        form = state.FORM
        if form == "cascade":
            raise TypeError("DirectForm<T,N> state needs a Cascade config")
        ctx.biquad(form, self.ba, self.F, self._clamp(), state.words, x, y, lanes=state.lanes, layout=layout)


class BiquadClamp(Biquad):
    """``BiquadClamp<C,T>``: offset ``u`` and limits at the summing junction
    (src/iir/biquad.rs:121-171).  Defaults: u=0, min=T::MIN, max=T::MAX."""

    def __init__(self, coeff: Biquad, u=None, min=None, max=None):
        super().__init__(coeff.ba, coeff.fmt)
        kind = self.kind
        if kind in _INT_INFO:
            info = np.iinfo(_INT_INFO[kind][0])
            lo, hi = info.min, info.max
        else:  # Clamp::MIN/MAX for floats are NEG_INFINITY / INFINITY (src/num.rs:5-31)
            lo, hi = -np.inf, np.inf
        self.u = 0 if u is None else u
        self.min = lo if min is None else min
        self.max = hi if max is None else max

    @classmethod
    def from_biquad(cls, coeff: Biquad) -> "BiquadClamp":
        return cls(coeff)

    def input_offset(self):
        return self.u / self.forward_gain()

    def set_input_offset(self, i):
        self.u = i * self.forward_gain()

    def _clamp(self):
        return [self.u, self.min, self.max]


class Cascade(_Proc):
    """``Cascade<[Biquad<C>; N]>`` on ``DirectForm<T, N>`` (src/iir/biquad.rs:321-364)."""

    def __init__(self, sections: Sequence[Biquad]):
        self.sections = list(sections)
        if not self.sections:
            raise ValueError("empty cascade")
        self.fmt = self.sections[0].fmt

    def _block(self, ctx, state, x, y, layout):
        if state.FORM != "cascade" or state.nsec != len(self.sections):
            raise TypeError("Cascade needs a DirectForm state with N = number of sections")
        ba = np.stack([s.ba for s in self.sections])
        F = self.fmt.F if isinstance(self.fmt, Q) else 0
        ctx.biquad("cascade", ba, F, None, state.words, x, y, lanes=state.lanes, layout=layout,
                   nsec=len(self.sections))


# ------------------------------------------------------------------ states
class _DfState(LaneState):
    FORM = ""

    def __init__(self, words, kind):
        super().__init__(words)
        self.kind = kind
        self.DTYPE = _INT_INFO[kind][0] if kind in _INT_INFO else _FLT[kind]

    @classmethod
    def _new(cls, nwords, lanes, kind, device, word_dtype=None):
        dt = word_dtype or (_INT_INFO[kind][0] if kind in _INT_INFO else _FLT[kind])
        return cls(LaneState._alloc(nwords, lanes, dt, device), kind)


class DirectForm1(_DfState):
    """``DirectForm1<T>`` = ``{x: [x0, x1], y: [[y0, y1]]}`` (biquad.rs:260-269, :321).
    words: [x[0], x[1], y[0][0], y[0][1]]."""

    FORM = "df1"

    @classmethod
    def default(cls, kind: str = "i32", lanes: int = 1, device=None):
        return cls._new(4, lanes, kind, device)

    @classmethod
    def from_xy(cls, x, y, kind="i32", device=None):
        s = cls.default(kind, 1, None)
        s.words[:, 0] = [x[0], x[1], y[0][0], y[0][1]]
        if device is not None:
            import torch
            s.words = torch.from_numpy(s.words).to(device)
        return s

    @property
    def x(self):
        return self.numpy()[0:2]

    @property
    def y(self):
        return self.numpy()[2:4]

    def set_y(self, v):
        """``set_y`` (biquad.rs:296-300): current and last output := v."""
        self.words[2:4] = v


class DirectForm2Transposed(_DfState):
    """``DirectForm2Transposed<T>`` = ``DirectForm<T,0,2>``: words [x[0], x[1]] (biquad.rs:407)."""

    FORM = "df2t"

    @classmethod
    def default(cls, kind: str = "f32", lanes: int = 1, device=None):
        return cls._new(2, lanes, kind, device)


class DirectForm1Wide(_DfState):
    """``DirectForm1Wide`` (biquad.rs:445-454): words (i32) [x0, x1, y0 lo, y0 hi, y1 lo, y1 hi]."""

    FORM = "df1wide"

    @classmethod
    def default(cls, lanes: int = 1, device=None):
        return cls._new(6, lanes, "i32", device)


class DirectForm1Dither(_DfState):
    """``DirectForm1Dither`` (biquad.rs:484-491): words (i32) [x0, x1, y0, y1, e]."""

    FORM = "df1dither"

    @classmethod
    def default(cls, lanes: int = 1, device=None):
        return cls._new(5, lanes, "i32", device)

    @classmethod
    def from_xye(cls, x, y, e, device=None):
        s = cls.default(1, None)
        s.words[:, 0] = [x[0], x[1], y[0][0], y[0][1], np.uint32(e).astype(np.int32)]
        if device is not None:
            import torch
            s.words = torch.from_numpy(s.words).to(device)
        return s


class DirectForm(_DfState):
    """``DirectForm<T, N>`` for a cascade of N sections (biquad.rs:260-269):
    words [x[0], x[1], y[0][0], y[0][1], ..., y[N-1][0], y[N-1][1]]."""

    FORM = "cascade"

    def __init__(self, words, kind):
        super().__init__(words, kind)
        self.nsec = (int(words.shape[0]) - 2) // 2

    @classmethod
    def default(cls, nsec: int, kind: str = "i32", lanes: int = 1, device=None):
        return cls._new(2 + 2 * nsec, lanes, kind, device)
