"""Host mirror of the NCO / lock-in pieces: ``cossin``, ``atan2``, ``Accu``,
``Lowpass<N>``, ``Lockin<C>`` and the four functions of the reference's Python
extension (``idsp._idsp``: cossin, atan2, sos, sos_clamp_wide -- src/py.rs).
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import FRAME_MAJOR, default_context
from .iir import Biquad, BiquadClamp, DirectForm1, DirectForm1Wide, Q32
from .process import LaneState, _Proc


def _dev_of(a):
    return a.device.index if isinstance(a, torch.Tensor) and a.is_cuda else None


def cossin(p):
    """``idsp.cossin(p: i32[n]) -> i32[n, 2]`` (src/py.rs:11-28, src/cossin.rs:14-67)."""
    if not isinstance(p, torch.Tensor):
        p = np.ascontiguousarray(p, np.int32).reshape(-1)
    ctx = default_context(_dev_of(p))
    out = ctx.cossin(p)
    if not (isinstance(p, torch.Tensor) and p.is_cuda):
        return out
    return out


def atan2(xy):
    """``idsp.atan2(xy: i32[n, 2]) -> i32[n]``: rows are (x, y) (src/py.rs:31-46)."""
    if not isinstance(xy, torch.Tensor):
        xy = np.ascontiguousarray(xy, np.int32)
    if xy.shape[-1] != 2:
        raise TypeError("shape")
    ctx = default_context(_dev_of(xy))
    return ctx.atan2(xy.reshape(-1))


def _round_sat_i32(v: float) -> int:
    """``f64::round() as i32`` (py.rs:99-101): half away from zero, saturating, NaN -> 0."""
    return Q32(0).from_float(v)


def sos(sos_rows, xy):
    """``idsp.sos(sos: f64[N,6], xy: i32[n])``: quantise each row
    ``[b0,b1,b2,a0,a1,a2]`` to ``Biquad<Q32<29>>``, fresh zero DF1 states, filter
    ``xy`` in place stage by stage (src/py.rs:50-74)."""
    rows = np.asarray(sos_rows, np.float64)
    if rows.ndim != 2 or rows.shape[1] != 6:
        raise TypeError("shape")
    dev = _dev_of(xy)
    for r in rows:
        bq = Biquad.from_ba6([r[0:3], r[3:6]], Q32(29))
        st = DirectForm1.default("i32", 1, None if dev is None else f"cuda:{dev}")
        bq.inplace(st, xy)


def sos_clamp_wide(sos_rows, xy):
    """``idsp.sos_clamp_wide(sos: f64[N,9], xy: i32[n])``: rows
    ``[b0,b1,b2,a0,a1,a2,u,min,max]`` on ``DirectForm1Wide`` (src/py.rs:80-108)."""
    rows = np.asarray(sos_rows, np.float64)
    if rows.ndim != 2 or rows.shape[1] != 9:
        raise TypeError("shape")
    dev = _dev_of(xy)
    for r in rows:
        bq = BiquadClamp(Biquad.from_ba6([r[0:3], r[3:6]], Q32(29)),
                         u=_round_sat_i32(r[6]), min=_round_sat_i32(r[7]), max=_round_sat_i32(r[8]))
        st = DirectForm1Wide.default(1, None if dev is None else f"cuda:{dev}")
        bq.inplace(st, xy)


class LowpassState(LaneState):
    """``LowpassState<N>``: words (i64) [state[0] .. state[N-1]] (src/lowpass.rs:17)."""

    DTYPE = np.int32

    @classmethod
    def default(cls, order: int, lanes: int = 1, device=None):
        return cls(LaneState._alloc(order, lanes, np.int64, device))


class Lowpass(_Proc):
    """``Lowpass<N>(pub [i32; N])`` (src/lowpass.rs:13, process :47-78)."""

    def __init__(self, k):
        self.k = [int(v) for v in k]
        if len(self.k) not in (1, 2):
            raise NotImplementedError("Lowpass order must be 1 or 2 (lowpass.rs:74-76)")

    def _block(self, ctx, state, x, y, layout):
        ctx.lowpass(self.k, state.words, x, y, lanes=state.lanes, layout=layout)


class Accu:
    """``Accu<Wrapping<i32>>`` per lane: ``state += step; yield state`` (src/accu.rs:15-38)."""

    def __init__(self, state, step):
        self.state, self.step = state, step


class LockinState(LaneState):
    """``[LowpassState<N>; 2]`` for the I and Q arms: words (i64) [I (N) | Q (N)]."""

    DTYPE = np.int32

    @classmethod
    def default(cls, order: int, lanes: int = 1, device=None):
        return cls(LaneState._alloc(2 * order, lanes, np.int64, device))


class Lockin:
    """``Lockin<Lowpass<N>>`` driven by (sample, phase) with the phase taken from a
    per-lane ``Accu`` (src/lockin.rs:30-39, src/accu.rs:34-37): output is
    ``Complex<i32>`` per sample, flat [..., 2]."""

    def __init__(self, lowpass: Lowpass):
        self.lowpass = lowpass

    def block(self, state: LockinState, accu: Accu, x, iq, layout: int = FRAME_MAJOR):
        ctx = default_context(_dev_of(x))
        ctx.lockin(self.lowpass.k, accu.state, accu.step, state.words, x, iq, lanes=state.lanes, layout=layout)

    def block_phase(self, state: LockinState, xp, iq, layout: int = FRAME_MAJOR):
        """``SplitProcess<(i32, Wrapping<i32>), Complex<i32>, [S; 2]>`` (src/lockin.rs:30-39): ``xp`` holds
        (sample, phase) pairs, the phase supplied by the caller (e.g. the PLL output, src/pll.rs:89-108)."""
        ctx = default_context(_dev_of(xp))
        ctx.lockin_phase(self.lowpass.k, state.words, xp, iq, lanes=state.lanes, layout=layout)

    def block_lo(self, state: LockinState, xlo, iq, layout: int = FRAME_MAJOR):
        """``SplitProcess<(X, Complex<U>), Complex<X>, [S; 2]>`` (src/lockin.rs:17-28) with X = i32,
        U = Q32<32>: ``xlo`` holds (sample, lo.re, lo.im) triples."""
        ctx = default_context(_dev_of(xlo))
        ctx.lockin_lo(self.lowpass.k, state.words, xlo, iq, lanes=state.lanes, layout=layout)


class PLLState(LaneState):
    """``PLLState`` (src/pll.rs:60-86) per lane as i32 words
    [clamp.x0, clamp.clamp, z0, y0, f0 lo, f0 hi, f lo, f hi, y]; all zero == ``default()``."""

    DTYPE = np.int32

    @classmethod
    def default(cls, lanes: int = 1, device=None):
        return cls(LaneState._alloc(9, lanes, np.int32, device))

    def phase(self):
        """``PLLState::phase()`` (src/pll.rs:76-78)"""
        return self.numpy()[8]

    def frequency(self):
        """``PLLState::frequency()`` = ``(f >> 32) as i32`` (src/pll.rs:81-83)"""
        return self.numpy()[7]


class PLL(_Proc):
    """``PLL { ba: [Q32<32>; 3] }`` (src/pll.rs:33-38) with raw coefficient bits."""

    def __init__(self, ba):
        self.ba = [int(v) for v in ba]
        if len(self.ba) != 3:
            raise ValueError("PLL.ba has three coefficients")

    @staticmethod
    def _q32(v: np.float32) -> int:
        """f32 -> Q32<32>: (v * 2^32).round() as i32, saturating (num_traits_impl.rs:30-45)"""
        s = np.float32(v) * np.float32(4294967296.0)
        s = np.float32(np.copysign(np.floor(np.abs(s) + np.float32(0.5)), s))  # round half away from zero
        if np.isnan(s):
            return 0
        return int(max(-(1 << 31), min((1 << 31) - 1, int(s))))

    @classmethod
    def from_zpk(cls, zero, pole, gain):
        """src/pll.rs:41-46 (f32 arithmetic)"""
        z, p, k = np.float32(zero), np.float32(pole), np.float32(gain)
        return cls([cls._q32(k), cls._q32(-k * z), cls._q32(-(np.float32(1.0) - p))])

    @classmethod
    def from_bandwidth(cls, bw, split=4.0):
        """src/pll.rs:51-57 (f32 arithmetic)"""
        bw, split = np.float32(bw), np.float32(split)
        a = bw * np.float32(2.0) * np.float32(np.pi)
        return cls.from_zpk(np.float32(1.0) - a / split, np.float32(1.0) - a * split, -a * a * split)

    def _block(self, ctx, state, x, y, layout):
        ctx.pll(self.ba, state.words, x, y, lanes=state.lanes, layout=layout)


class FmDiscState(LaneState):
    """State of the fused FM discriminator graph per lane, i32 words
    [has_prev, prev.re, prev.im, x1, x2, y1, y2] (``Option<Complex<Q32<32>>>`` + ``DirectForm1<i32>``)."""

    DTYPE = np.int32

    @classmethod
    def default(cls, lanes: int = 1, device=None):
        return cls(LaneState._alloc(7, lanes, np.int32, device))


class FmDiscriminator(_Proc):
    """``(FmDiscriminator{carrier} * Biquad<Q32<F>>).minor()`` of examples/fm_disc.rs:26-48:
    X = Complex<Q32<32>> as (re, im) i32 pairs -> Y = i32."""

    def __init__(self, carrier: int, deemph):
        self.carrier = int(carrier)
        self.deemph = deemph  # iir.Biquad with an i32 Q format

    def widths(self):
        return 2, 1

    def _block(self, ctx, state, x, y, layout):
        ctx.fm_disc(self.carrier, self.deemph.ba, self.deemph.fmt.F, state.words, x, y, lanes=state.lanes, layout=layout)
