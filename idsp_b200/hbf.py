"""Host mirror of ``idsp::hbf`` (half-band FIR decimators / interpolators).

  EvenSymmetric<[C;M]> as /2 decimator  on HbfDec   src/hbf.rs:155-192
  EvenSymmetric<[C;M]> as x2 interpolator on HbfInt src/hbf.rs:207-236
  HBF_TAPS                                          src/hbf.rs:308-349
  HBF_DEC_CASCADE / HbfDec2..32                     src/hbf.rs:363-421
  HBF_INT_CASCADE / HbfInt2..32                     src/hbf.rs:454-512
  OddSymmetric / EvenSymmetric / Odd/EvenAntiSymmetric single-rate FIRs  src/hbf.rs:70-138
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .process import LaneState, _Proc


def _taps(idx: int) -> np.ndarray:
    m = C.c_int(0)
    p = _lib.lib().idsp_hbf_taps(idx, C.byref(m))
    return np.ctypeslib.as_array(p, shape=(m.value,)).copy()


def hbf_taps():
    """``HBF_TAPS``: tuple of 5 tap arrays, index 0 = lowest rate (hbf.rs:308-349)."""
    return tuple(_taps(i) for i in range(5))


def hbf_taps_98():
    """``HBF_TAPS_98``: the 98 dB tap set, index 0 = lowest rate (hbf.rs:258-292)."""
    out = []
    for i in range(5):
        m = C.c_int(0)
        p = _lib.lib().idsp_hbf_taps_98(i, C.byref(m))
        out.append(np.ctypeslib.as_array(p, shape=(m.value,)).copy())
    return tuple(out)


HBF_PASSBAND = 0.4  # hbf.rs:352
HBF_CASCADE_BLOCK = 1 << 5  # hbf.rs:357 (CPU heuristic; the GPU tiles differently)


def hbf_dec_response_length(depth: int) -> int:
    """hbf.rs:424-448"""
    assert depth <= 5
    n = 0
    for i in range(depth - 1, -1, -1):
        n //= 2
        n += 2 * len(_taps(i)) - 1
    return n


def hbf_int_response_length(depth: int) -> int:
    """hbf.rs:515-539"""
    assert depth <= 5
    n = 0
    for i in range(depth):
        n += 2 * len(_taps(i)) - 1
        n *= 2
    return n


class HbfDec(LaneState):
    """``HbfDec<[T;N]>`` (hbf.rs:142-153): words [even history (M-1) | odd history (2M-1)],
    oldest first.  The reference's N (block capacity) is not a property of the GPU state."""

    DTYPE = np.float32

    def __init__(self, words, M):
        super().__init__(words)
        self.M = M

    @classmethod
    def default(cls, M: int, lanes: int = 1, device=None):
        return cls(LaneState._alloc(3 * M - 2, lanes, np.float32, device), M)


class HbfInt(LaneState):
    """``HbfInt<[T;N]>`` (hbf.rs:196-205): words [x history (2M-1)]."""

    DTYPE = np.float32

    def __init__(self, words, M):
        super().__init__(words)
        self.M = M

    @classmethod
    def default(cls, M: int, lanes: int = 1, device=None):
        return cls(LaneState._alloc(2 * M - 1, lanes, np.float32, device), M)


class FirState(LaneState):
    DTYPE = np.float32

    @classmethod
    def default(cls, nwords: int, lanes: int = 1, device=None):
        return cls(LaneState._alloc(nwords, lanes, np.float32, device))


class _Fir(_Proc):
    ODD = False
    SYM = True

    def __init__(self, taps):
        self.taps = np.asarray(taps, np.float32).reshape(-1)

    @property
    def M(self):
        return int(self.taps.size)

    def len(self) -> int:
        """``LEN``: response length minus one (hbf.rs:78)."""
        return 2 * self.M - 1 + int(self.ODD)

    def widths(self):
        return 1, 1

    def _block(self, ctx, state, x, y, layout):
        # the kernels size the delay lines from the taps (3M-2 / 2M-1 / 2M-1+odd words): a state built for
        # another M would be read and written past its end on the device
        if isinstance(state, (HbfDec, HbfInt)):
            if state.M != self.M:
                raise TypeError(f"state was built for M = {state.M} taps, the filter has M = {self.M}")
            if not (self.SYM and not self.ODD):
                raise TypeError("HbfDec / HbfInt states belong to EvenSymmetric (hbf.rs:155-236)")
        if isinstance(state, HbfDec):
            ctx.hbf_dec(self.taps, state.words, x, y, lanes=state.lanes, layout=layout)
        elif isinstance(state, HbfInt):
            ctx.hbf_int(self.taps, state.words, x, y, lanes=state.lanes, layout=layout)
        else:
            if state.words.shape[0] != self.len():
                raise TypeError(f"FIR state has {state.words.shape[0]} words, the filter needs LEN = {self.len()}")
            ctx.fir(self.taps, self.ODD, self.SYM, state.words, x, y, lanes=state.lanes, layout=layout)

    def block(self, state, x, y, layout=0):
        # rate depends on the state type, like the reference's three impls on EvenSymmetric
        from .process import _numel
        wi, wo = (2, 1) if isinstance(state, HbfDec) else (1, 2) if isinstance(state, HbfInt) else (1, 1)
        if _numel(x) * wo != _numel(y) * wi:
            raise ValueError("block: x and y lengths do not match")
        if _numel(x) % (state.lanes * wi):
            raise ValueError("block: length is not a whole number of frames")
        self._block(self._ctx(state), state, x, y, layout)


class OddSymmetric(_Fir):
    ODD, SYM = True, True


class EvenSymmetric(_Fir):
    ODD, SYM = False, True


class OddAntiSymmetric(_Fir):
    ODD, SYM = True, False


class EvenAntiSymmetric(_Fir):
    ODD, SYM = False, False


class HbfCascadeState(LaneState):
    DTYPE = np.float32

    def __init__(self, words, log2_rate):
        super().__init__(words)
        self.log2_rate = log2_rate


def _dec_state(k):
    def default(lanes: int = 1, device=None):
        n = int(_lib.lib().idsp_hbf_dec_state_words(k))
        return HbfCascadeState(LaneState._alloc(n, lanes, np.float32, device), k)
    return default


def _int_state(k):
    def default(lanes: int = 1, device=None):
        n = int(_lib.lib().idsp_hbf_int_state_words(k))
        return HbfCascadeState(LaneState._alloc(n, lanes, np.float32, device), k)
    return default


# ``HbfDec2 .. HbfDec32`` / ``HbfInt2 .. HbfInt32`` default-state constructors (hbf.rs:363-383, 454-474)
HbfDec2, HbfDec4, HbfDec8, HbfDec16, HbfDec32 = (_dec_state(k) for k in range(1, 6))
HbfInt2, HbfInt4, HbfInt8, HbfInt16, HbfInt32 = (_int_state(k) for k in range(1, 6))


def _cascade_words(decimate: bool, taps) -> int:
    return sum((3 * len(t) - 2) if decimate else (2 * len(t) - 1) for t in taps)


class HbfDecCascade(_Proc):
    """``HBF_DEC_CASCADE`` truncated to depth k (``.inner().1`` ... in the reference,
    hbf.rs:385-421): X = [f32; 2^k] -> Y = f32, stages TAPS[k-1] -> TAPS[0].
    ``taps``: another tuple of half-band tap sets in the reference's order (index 0 = lowest rate),
    e.g. ``hbf_taps_98()[:k]``; the default is ``HBF_TAPS``."""

    def __init__(self, log2_rate: int, taps=None):
        if not 1 <= log2_rate <= 5:
            raise ValueError("log2_rate must be 1..5")
        self.k = log2_rate
        self.taps = None if taps is None else [np.asarray(t, np.float32).reshape(-1) for t in taps]
        if self.taps is not None and len(self.taps) != self.k:
            raise ValueError("one tap set per stage")

    def widths(self):
        return 1 << self.k, 1

    def state(self, lanes: int = 1, device=None):
        """default (zero) state for this cascade"""
        if self.taps is None:
            return _dec_state(self.k)(lanes, device)
        return HbfCascadeState(LaneState._alloc(_cascade_words(True, self.taps), lanes, np.float32, device), self.k)

    def _block(self, ctx, state, x, y, layout):
        if state.log2_rate != self.k:
            raise TypeError("state depth does not match cascade depth")
        if self.taps is None:
            ctx.hbf_dec_cascade(self.k, state.words, x, y, lanes=state.lanes, layout=layout)
        else:
            ctx.hbf_cascade_taps(True, self.taps, state.words, x, y, lanes=state.lanes, layout=layout)


class HbfIntCascade(_Proc):
    """``HBF_INT_CASCADE`` truncated to depth k (hbf.rs:476-512): X = f32 -> Y = [f32; 2^k]."""

    def __init__(self, log2_rate: int, taps=None):
        if not 1 <= log2_rate <= 5:
            raise ValueError("log2_rate must be 1..5")
        self.k = log2_rate
        self.taps = None if taps is None else [np.asarray(t, np.float32).reshape(-1) for t in taps]
        if self.taps is not None and len(self.taps) != self.k:
            raise ValueError("one tap set per stage")

    def widths(self):
        return 1, 1 << self.k

    def state(self, lanes: int = 1, device=None):
        if self.taps is None:
            return _int_state(self.k)(lanes, device)
        return HbfCascadeState(LaneState._alloc(_cascade_words(False, self.taps), lanes, np.float32, device), self.k)

    def _block(self, ctx, state, x, y, layout):
        if state.log2_rate != self.k:
            raise TypeError("state depth does not match cascade depth")
        if self.taps is None:
            ctx.hbf_int_cascade(self.k, state.words, x, y, lanes=state.lanes, layout=layout)
        else:
            ctx.hbf_cascade_taps(False, self.taps, state.words, x, y, lanes=state.lanes, layout=layout)
