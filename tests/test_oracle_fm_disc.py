"""Pins the FM-discriminator graph of the CPU oracle with the reference
example's own self-test (examples/fm_disc.rs:126-158: the recovered message correlates > 0.999 with
the low-passed modulation, gain within 5 %, rms < 5e-4) and with a step-by-step composition of the
already pinned oracle pieces (cossin / atan2 / DF1).  CPU only."""
import numpy as np

from idsp_b200.coefficients import Filter
from idsp_b200.iir import Biquad, Q32

TAU = np.float32(2 * np.pi)


def _fm_signal(oracle, carrier, deviation, message_freq, n):
    """examples/fm_disc.rs:57-75 (f32 fixture)"""
    i = np.arange(n, dtype=np.float32)
    msg = np.sin(TAU * np.float32(message_freq) * i).astype(np.float32)
    inc = (np.int64(np.int32(np.uint32(carrier))) + (np.float32(deviation) * msg).astype(np.int32).astype(np.int64))
    phase = (np.cumsum(inc) & 0xffffffff).astype(np.uint32).view(np.int32)
    return oracle.cossin(phase), msg


def _lowpass_f32(oracle, cutoff, x):
    """examples/fm_disc.rs:92-100: Biquad::<f32> DF1 lowpass"""
    ba = Biquad.from_ba6(Filter().critical_frequency(cutoff).lowpass(), "f32").ba
    st = np.zeros(4, np.float32)
    return oracle.biquad_df1("f32", ba, 0, None, st, np.asarray(x, np.float32))


def run_fm_disc(oracle, n=4096):
    """examples/fm_disc.rs:126-140"""
    carrier, deviation, message_freq = 0x19341234, 0x04500000, 0.004
    scale = TAU / np.float32(4294967296.0)
    x, msg = _fm_signal(oracle, carrier, deviation, message_freq, n)
    ba = Biquad.from_ba6(Filter().critical_frequency(0.02).lowpass(), Q32(30)).ba
    st = np.zeros((7, 1), np.int32)
    y = oracle.fm_disc_lanes(carrier, ba, 30, st, x.reshape(-1), 1).astype(np.float32) * scale
    m = _lowpass_f32(oracle, 0.02, np.float32(deviation) * scale * msg)
    return y[1024:], m[1024:], x, ba


def test_tracks_known_modulation(oracle):
    """examples/fm_disc.rs:149-158"""
    y, m, _, _ = run_fm_disc(oracle)
    y64, m64 = y.astype(np.float64), m.astype(np.float64)
    gain = (y64 * m64).sum() / (m64 * m64).sum()
    rms = np.sqrt(((y64 - gain * m64) ** 2).sum()) / len(y)
    corr = (y64 * m64).sum() / (np.sqrt((y64 * y64).sum()) * np.sqrt((m64 * m64).sum()))
    assert corr > 0.999 and 0.95 < gain < 1.05 and rms < 5e-4


def test_graph_equals_composition_of_pinned_pieces(oracle):
    """fused graph == x[n] * conj(x[n-1]) with big-int products, oracle.atan2, oracle DF1 (each pinned
    by the reference's own tests), first output from d = 0; state words carry prev and the DF1 state"""
    _, _, x, ba = run_fm_disc(oracle, 600)
    carrier = 0x19341234
    st = np.zeros((7, 1), np.int32)
    got = oracle.fm_disc_lanes(carrier, ba, 30, st, x.reshape(-1), 1)
    d = [0]
    for n in range(1, len(x)):
        xr, xi, pr, pi = (int(v) for v in (x[n, 0], x[n, 1], x[n - 1, 0], x[n - 1, 1]))
        re, im = (xr * pr + xi * pi) >> 32, (xi * pr - xr * pi) >> 32
        a = int(oracle.atan2(np.array([[re, im]], np.int32))[0])
        d.append((a - carrier + (1 << 31)) % (1 << 32) - (1 << 31))
    s = np.zeros(4, np.int32)
    want = oracle.biquad_df1("i32", ba, 30, None, s, np.array(d, np.int32))
    assert np.array_equal(got, want)
    assert [int(v) for v in st[:, 0]] == [1, int(x[-1, 0]), int(x[-1, 1])] + [int(v) for v in s]
