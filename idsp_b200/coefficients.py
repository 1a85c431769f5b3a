"""Host-side biquad coefficient builder: mirror of ``idsp::iir::coefficients::Filter``
(src/iir/coefficients.rs:17-40, 111-527; audio-EQ-cookbook formulas).  Host math only --
it feeds raw coefficients to the C ABI, it is not part of the device hot path.

Two implementations that must agree: the f64 formulas below (pure Python, libm through
``math``) and the C ABI builders ``idsp_filter_build_{f64,f32}`` (``idsp_b200/csrc/coeff.cu``),
which also carry the reference's f32 flavour (``Filter<f32>``: every intermediate rounded to
f32).  ``Filter(dtype="f32")`` routes through the C ABI.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

# coefficients::Type (coefficients.rs:44-66), the order of the C ABI's idsp_filter_type_t
TYPES = ("lowpass", "highpass", "bandpass", "allpass", "notch", "peaking", "lowshelf", "highshelf", "iho")
_SHAPES = {"q": 0, "bandwidth": 1, "slope": 2}


class FilterError(ValueError):
    """``iir::Error`` (src/iir/error.rs:5-16): the message is ``Variant(field)``"""


@dataclass
class Shape:
    kind: str = "q"  # q | bandwidth | slope
    value: float = 1.0 / math.sqrt(2.0)  # Shape::default() = Q(1/sqrt(2)) (coefficients.rs:18-22)


@dataclass
class Filter:
    frequency: float = 0.0
    gain: float = 1.0
    shelf: float = 1.0
    shape: Shape = field(default_factory=Shape)
    dtype: str = "f64"  # the reference's Filter<f64> or Filter<f32>

    # ---- C ABI route (both widths; the only route for f32)
    def _c_struct(self):
        from . import _lib

        st = (_lib.FilterF32 if self.dtype == "f32" else _lib.FilterF64)()
        st.frequency, st.gain, st.shelf = self.frequency, self.gain, self.shelf
        st.shape_kind, st.shape = _SHAPES[self.shape.kind], self.shape.value
        return st

    def build_c(self, typ: str):
        """``Filter::build(typ)`` through ``idsp_filter_build_{f64,f32}``"""
        from . import _lib

        L = _lib.lib()
        st = self._c_struct()
        out = ((C.c_float if self.dtype == "f32" else C.c_double) * 6)()
        fn = L.idsp_filter_build_f32 if self.dtype == "f32" else L.idsp_filter_build_f64
        _lib.check(fn(C.byref(st), TYPES.index(typ), out))
        v = list(out)
        return [v[0:3], v[3:6]]

    def validate(self):
        """``Filter::validate`` (coefficients.rs:241-265); raises FilterError(``Variant(field)``)"""
        f, g, a = self.frequency, self.gain, self.shelf
        fin = math.isfinite
        if not fin(f):
            raise FilterError("NonFinite(frequency)")
        if f < 0.0 or f > math.pi:
            raise FilterError("OutOfRange(frequency)")
        if not fin(g) or g <= 0.0:
            raise FilterError("NonPositive(gain)")
        if not fin(a) or a <= 0.0:
            raise FilterError("NonPositive(shelf)")
        k, v = self.shape.kind, self.shape.value
        name = {"q": "q", "bandwidth": "bandwidth", "slope": "slope"}[k]
        if not fin(v):
            raise FilterError(f"NonFinite({name})")
        if k != "bandwidth" and v <= 0.0:
            raise FilterError(f"NonPositive({name})")

    def build(self, typ: str):
        """``Filter::build(typ)`` (coefficients.rs:466-479): ``[[b0,b1,b2],[a0,a1,a2]]``"""
        if typ not in TYPES:
            raise ValueError(f"type must be one of {TYPES}")
        if self.dtype == "f32":
            return self.build_c(typ)
        return getattr(self, typ)()

    def try_build(self, typ: str):
        self.validate()
        return self.build(typ)

    def build_biquad(self, typ: str, fmt):
        """``Filter::build_biquad::<C>(typ)`` (coefficients.rs:481-487)"""
        from .iir import Biquad

        return Biquad.from_ba6(self.build(typ), fmt, src=self.dtype)

    def try_build_biquad(self, typ: str, fmt):
        self.validate()
        return self.build_biquad(typ, fmt)

    def build_clamped(self, typ: str, fmt):
        """``Filter::build_clamped`` (coefficients.rs:499-505): default clamp (u = 0, MIN, MAX)"""
        from .iir import BiquadClamp

        return BiquadClamp(self.build_biquad(typ, fmt))

    # builder methods (coefficients.rs:111-238)
    def set_frequency(self, critical_frequency, sample_frequency):
        return self.critical_frequency(critical_frequency / sample_frequency)

    def critical_frequency(self, f0):
        return self.angular_critical_frequency(math.tau * f0)

    def angular_critical_frequency(self, w0):
        self.frequency = w0
        return self

    def set_gain(self, k):
        self.gain = k
        return self

    def gain_db(self, k_db):
        return self.set_gain(10.0 ** (k_db / 20.0))

    def set_shelf(self, a):
        self.shelf = a
        return self

    def shelf_db(self, a_db):
        return self.set_shelf(10.0 ** (a_db / 20.0))

    def inverse_q(self, qi):
        return self.q(1.0 / qi)

    def q(self, q):
        self.shape = Shape("q", q)
        return self

    def bandwidth(self, bw):
        self.shape = Shape("bandwidth", bw)
        return self

    def shelf_slope(self, s):
        self.shape = Shape("slope", s)
        return self

    # coefficients.rs:266-283
    def _qi(self):
        k, v = self.shape.kind, self.shape.value
        if k == "q":
            return 1.0 / v
        if k == "bandwidth":
            return 2.0 * math.sinh(math.log(2.0) / 2.0 * v * self.frequency / math.sin(self.frequency))
        return math.sqrt((self.shelf + 1.0 / self.shelf) * (1.0 / v - 1.0) + 2.0)

    def _fcos_alpha(self):
        fsin, fcos = math.sin(self.frequency), math.cos(self.frequency)
        return fcos, 0.5 * fsin * self._qi()

    # coefficients.rs:302-470
    def lowpass(self):
        fcos, alpha = self._fcos_alpha()
        b = self.gain * 0.5 * (1.0 - fcos)
        return [[b, 2.0 * b, b], [1.0 + alpha, -2.0 * fcos, 1.0 - alpha]]

    def highpass(self):
        fcos, alpha = self._fcos_alpha()
        b = self.gain * 0.5 * (1.0 + fcos)
        return [[b, -2.0 * b, b], [1.0 + alpha, -2.0 * fcos, 1.0 - alpha]]

    def bandpass(self):
        fcos, alpha = self._fcos_alpha()
        b = self.gain * alpha
        return [[b, 0.0, -b], [1.0 + alpha, -2.0 * fcos, 1.0 - alpha]]

    def notch(self):
        fcos, alpha = self._fcos_alpha()
        f2 = -2.0 * fcos
        return [[self.gain, f2 * self.gain, self.gain], [1.0 + alpha, f2, 1.0 - alpha]]

    def allpass(self):
        fcos, alpha = self._fcos_alpha()
        f2 = -2.0 * fcos
        return [[(1.0 - alpha) * self.gain, f2 * self.gain, (1.0 + alpha) * self.gain],
                [1.0 + alpha, f2, 1.0 - alpha]]

    def peaking(self):
        fcos, alpha = self._fcos_alpha()
        s = math.sqrt(self.shelf)
        f2 = -2.0 * fcos
        return [[(1.0 + alpha * s) * self.gain, f2 * self.gain, (1.0 - alpha * s) * self.gain],
                [1.0 + alpha / s, f2, 1.0 - alpha / s]]

    def lowshelf(self):
        fcos, alpha = self._fcos_alpha()
        s = math.sqrt(self.shelf)
        tsa = 2.0 * math.sqrt(s) * alpha
        sp1, sm1 = s + 1.0, s - 1.0
        return [[s * self.gain * (sp1 - sm1 * fcos + tsa), 2.0 * s * self.gain * (sm1 - sp1 * fcos),
                 s * self.gain * (sp1 - sm1 * fcos - tsa)],
                [sp1 + sm1 * fcos + tsa, -2.0 * (sm1 + sp1 * fcos), sp1 + sm1 * fcos - tsa]]

    def highshelf(self):
        fcos, alpha = self._fcos_alpha()
        s = math.sqrt(self.shelf)
        tsa = 2.0 * math.sqrt(s) * alpha
        sp1, sm1 = s + 1.0, s - 1.0
        return [[s * self.gain * (sp1 + sm1 * fcos + tsa), -2.0 * s * self.gain * (sm1 + sp1 * fcos),
                 s * self.gain * (sp1 + sm1 * fcos - tsa)],
                [sp1 - sm1 * fcos + tsa, 2.0 * (sm1 - sp1 * fcos), sp1 - sm1 * fcos - tsa]]

    def iho(self):
        """I/HO: notch, integrating below, flat ``shelf`` gain above (coefficients.rs:451-464)"""
        fcos, alpha = self._fcos_alpha()
        fsin = 0.5 * math.sin(self.frequency)
        a = (1.0 + fcos) / (2.0 * self.shelf)
        return [[self.gain * (1.0 + alpha), -2.0 * self.gain * fcos, self.gain * (1.0 - alpha)],
                [a + fsin, -2.0 * a, a - fsin]]


@dataclass
class WebAudio:
    """``coefficients::WebAudio`` (coefficients.rs:68-86, 529-560): WebAudio-style parametrisation"""

    typ: str = "lowpass"
    frequency_hz: float = 350.0
    sample_rate_hz: float = 48e3
    detune_cents: float = 0.0
    q: float = 1.0
    gain_db: float = 0.0

    def filter(self) -> Filter:
        f = Filter()
        f.set_frequency(self.frequency_hz * 2.0 ** (self.detune_cents / 1200.0), self.sample_rate_hz)
        f.q(self.q)
        if self.typ in ("peaking", "lowshelf", "highshelf"):
            f.shelf_db(self.gain_db)
        return f

    def build(self):
        return self.filter().build(self.typ)

