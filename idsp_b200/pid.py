"""Host mirror of ``idsp::iir::pid::Builder`` (src/iir/pid.rs:27-317) on top of the C ABI builders
(``idsp_pid_build_{f64,f32}``, ``idsp_b200/csrc/coeff.cu``): PID action gains and gain limits ->
``Biquad`` coefficients.  Host math only."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _lib
from .iir import Biquad, Q, _FLT, _INT_INFO

# pid::Order (pid.rs:13-24) and pid::Action (pid.rs:58-72)
ORDER = {"P": 2, "I": 1, "I2": 0}
ACTION = {"I2": 0, "I": 1, "P": 2, "D": 3, "D2": 4}


class PidError(ValueError):
    """``iir::Error`` raised by ``validate`` / ``try_build``: ``Variant(field)``"""


@dataclass
class Builder:
    """``pid::Builder<T>``; Default = order I, zero gains, +inf limits (pid.rs:49-57)."""

    order: str = "I"
    gains: List[float] = field(default_factory=lambda: [0.0] * 5)
    limits: List[float] = field(default_factory=lambda: [float("inf")] * 5)
    dtype: str = "f64"  # Builder<f64> or Builder<f32>

    def set_order(self, order: str) -> "Builder":
        self.order = order
        return self

    def gain(self, action: str, gain: float) -> "Builder":
        self.gains[ACTION[action]] = gain
        return self

    def limit(self, action: str, limit: float) -> "Builder":
        self.limits[ACTION[action]] = limit
        return self

    def kp(self, g): return self.gain("P", g)
    def ki(self, g): return self.gain("I", g)
    def ki2(self, g): return self.gain("I2", g)
    def kd(self, g): return self.gain("D", g)
    def kd2(self, g): return self.gain("D2", g)
    def limit_i(self, v): return self.limit("I", v)
    def limit_i2(self, v): return self.limit("I2", v)
    def limit_d(self, v): return self.limit("D", v)
    def limit_d2(self, v): return self.limit("D2", v)

    def _c(self):
        st = (_lib.PidF32 if self.dtype == "f32" else _lib.PidF64)()
        st.order = ORDER[self.order]
        for i in range(5):
            st.gain[i], st.limit[i] = self.gains[i], self.limits[i]
        return st

    def validate(self, period: float) -> None:
        """pid.rs:193-222"""
        L = _lib.lib()
        fn = L.idsp_pid_validate_f32 if self.dtype == "f32" else L.idsp_pid_validate_f64
        if fn(C.byref(self._c()), period) != 0:
            raise PidError(L.idsp_b200_last_error().decode())

    def build(self, period: float, fmt) -> Biquad:
        """``Build<Biquad<C>>::build(&period)`` (pid.rs:236-317); fmt = a ``Q`` format or 'f32' / 'f64'"""
        L = _lib.lib()
        kind = fmt.kind if isinstance(fmt, Q) else fmt
        dt = _INT_INFO[kind][0] if kind in _INT_INFO else _FLT[kind]
        out = np.zeros(5, dt)
        fn = L.idsp_pid_build_f32 if self.dtype == "f32" else L.idsp_pid_build_f64
        _lib.check(fn(C.byref(self._c()), period, _lib.KIND_CODE[kind], fmt.F if isinstance(fmt, Q) else 0,
                      out.ctypes.data_as(C.c_void_p)))
        return Biquad(out, fmt)

    def try_build(self, period: float, fmt) -> Biquad:
        self.validate(period)
        return self.build(period, fmt)
