"""Host mirror of ``idsp::cic``.

  Cic<T, N, M>                         src/cic.rs:13-29 (state), :31-147 (new / gain / ...)
  Process<T, Option<T>> (decimator)    src/cic.rs:176-200
  Process<Option<T>, T> (interpolator) src/cic.rs:149-172
  .decimate() / .interpolate()         dsp-process/src/adapters.rs:154-222 / :27-35

In the reference ``Cic`` is a stateful ``Process``; here the configuration (N, M, rate) and the per-lane
state (``CicState``) are split like every other filter of this package, and ``Decimator(Cic(..))`` /
``Interpolator(Cic(..))`` are the chunk adapters a ``Lanes`` runs on ``[T; rate+1]`` frames.
"""
from __future__ import annotations

import numpy as np

from .process import LaneState, _Proc


class CicState(LaneState):
    """Per-lane ``Cic`` fields as SoA words of T: [index, zoh, combs[N][M], integrators[N]];
    all zero == ``Cic::new(rate)`` (src/cic.rs:39-47)."""

    @classmethod
    def default(cls, N: int, M: int = 1, dtype="i64", lanes: int = 1, device=None):
        st = cls(LaneState._alloc(2 + N * M + N, lanes, {"i32": np.int32, "i64": np.int64}[dtype], device))
        st.DTYPE = {"i32": np.int32, "i64": np.int64}[dtype]
        return st


class Cic:
    """``Cic::<T, N, M>::new(rate)`` configuration: order N (1..6), comb delay M (1..3),
    rate = fast/slow - 1."""

    def __init__(self, N: int, M: int = 1, rate: int = 0):
        if not (1 <= N <= 6 and 1 <= M <= 3):
            raise ValueError("Cic lanes: N must be 1..6 and M 1..3")
        if not 0 <= rate < (1 << 32):
            raise ValueError("rate is a u32")
        self.N, self.M, self.rate = N, M, rate

    def order(self):
        return self.N

    def comb_delay(self):
        return self.M

    def gain(self) -> int:
        """src/cic.rs:99-101"""
        return (self.M * (self.rate + 1)) ** self.N

    def gain_log2(self) -> int:
        """src/cic.rs:107-109"""
        return (self.M * self.rate + self.M - 1).bit_length() * self.N

    def response_length(self) -> int:
        """src/cic.rs:112-114"""
        return self.rate * self.N

    def decimate(self):
        return Decimator(self)

    def interpolate(self):
        return Interpolator(self)


class Decimator(_Proc):
    """``Decimator(Cic)``: X = [T; rate+1] -> Y = T (the value of the frame's tick)."""

    def __init__(self, cic: Cic):
        self.cic = cic

    def widths(self):
        return self.cic.rate + 1, 1

    def _block(self, ctx, state, x, y, layout):
        c = self.cic
        if state.words.shape[0] != 2 + c.N * c.M + c.N:
            raise TypeError("CicState does not match Cic<N, M>")
        ctx.cic(True, c.N, c.M, c.rate, state.words, x, y, lanes=state.lanes, layout=layout)


class Interpolator(_Proc):
    """``Interpolator(Cic)``: X = T -> Y = [T; rate+1]."""

    def __init__(self, cic: Cic):
        self.cic = cic

    def widths(self):
        return 1, self.cic.rate + 1

    def _block(self, ctx, state, x, y, layout):
        c = self.cic
        if state.words.shape[0] != 2 + c.N * c.M + c.N:
            raise TypeError("CicState does not match Cic<N, M>")
        ctx.cic(False, c.N, c.M, c.rate, state.words, x, y, lanes=state.lanes, layout=layout)
