"""Host-side biquad coefficient builder: mirror of ``idsp::iir::coefficients::Filter``
(src/iir/coefficients.rs:17-40, 111-527; audio-EQ-cookbook formulas).  Pure host
math in f64 -- it feeds raw coefficients to the C ABI, it is not part of the
device hot path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field


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
