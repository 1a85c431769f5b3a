"""idsp_b200 -- B200-native multi-lane engine for the quartiq/idsp filter hot path.

The arithmetic lives in hand-written sm_100a CUDA kernels behind the C ABI of
``include/idsp_b200.h`` (``libidsp_b200.so``); this package is the host-side
mirror of the reference's operator surface (``dsp_process::{SplitProcess, Split,
Lanes, View}``, ``idsp::iir``, ``idsp::hbf``, ``cossin``/``atan2``,
``Lowpass``/``Lockin``) plus the four functions of its Python extension.
There is no CPU implementation: importing works without a GPU, calling needs one.
"""
from .engine import FRAME_MAJOR, LANE_MAJOR, Context, default_context  # noqa: F401
from .process import FrameMajor, LaneMajor, Lanes, Split, View, ViewMut  # noqa: F401
from .iir import (  # noqa: F401
    Biquad, BiquadClamp, Cascade, DirectForm, DirectForm1, DirectForm1Dither, DirectForm1Wide,
    DirectForm2Transposed, Q, Q8, Q16, Q32, Q64,
)
from .hbf import (  # noqa: F401
    EvenAntiSymmetric, EvenSymmetric, HbfDec, HbfDec2, HbfDec4, HbfDec8, HbfDec16, HbfDec32,
    HbfDecCascade, HbfInt, HbfInt2, HbfInt4, HbfInt8, HbfInt16, HbfInt32, HbfIntCascade,
    OddAntiSymmetric, OddSymmetric, hbf_dec_response_length, hbf_int_response_length, hbf_taps, hbf_taps_98,
)
from .nco import (  # noqa: F401
    PLL, Accu, FmDiscriminator, FmDiscState, Lockin, LockinState, Lowpass, LowpassState, PLLState, atan2, cossin, sos, sos_clamp_wide,
)
from .coefficients import Filter, FilterError, WebAudio  # noqa: F401
from . import pid  # noqa: F401
from .cic import Cic, CicState, Decimator, Interpolator  # noqa: F401

__version__ = "0.1.0"
