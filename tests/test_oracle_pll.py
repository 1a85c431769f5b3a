"""Pins the PLL part of the CPU oracle with the reference's own tests
(src/pll.rs:117-149: convergence of a wide and a narrow loop on an Accu phase ramp) and against an
independent big-integer Python model of src/pll.rs:88-108.  CPU only."""
import numpy as np


def _wrap(v, bits):
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def _py_pll(ba, x):
    """plain-int restatement written from the prose of src/pll.rs / src/unwrap.rs"""
    x0 = clamp = z0p = y0p = f0 = f = y = 0
    out = []
    for xv in x:
        y = _wrap(y + (f >> 32), 32)
        t = _wrap(int(xv) + y, 32)
        delta = _wrap(t - x0, 32)
        a, b = delta >= 0, t >= x0
        wrap = (a > b) - (a < b)
        x0 = t
        c = clamp + wrap
        clamp = (c > 0) - (c < 0)
        o = -(1 << 31) if clamp < 0 else ((1 << 31) - 1 if clamp > 0 else t)
        z0 = o >> 1
        y0 = _wrap(z0 + z0p, 32)
        z0p = z0
        acc = ba[0] * y0 + ba[1] * y0p + ba[2] * _wrap(f0 >> 32, 32) + ((ba[2] * (f0 & 0xffffffff)) >> 32)
        f0 = _wrap(f0 + acc, 64)
        y0p = y0
        f = _wrap(f + f0, 64)
        out.append(y)
    return out, (x0, clamp, z0p, y0p, f0, f, y)


def _ramp(step, n):
    """Accu::<W<i32>>::new(W(0), W(step)).take(n) (src/accu.rs:29-38: pre-increment)"""
    return ((np.arange(1, n + 1, dtype=np.uint64) * np.uint64(step)) & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32)


def test_converge_pll(oracle):
    """src/pll.rs:117-132"""
    ba = oracle.pll_from_bandwidth(5e-2, 4.0)
    n, step = 1 << 9, 0x71f63049
    x = _ramp(step, n)
    st = np.zeros((9, 1), np.int32)
    y = oracle.pll_lanes(ba, st, x, 1)
    # replay to read the frequency after every step
    st2 = np.zeros((9, 1), np.int32)
    for i in range(n):
        yi = oracle.pll_lanes(ba, st2, x[i:i + 1], 1)
        assert yi[0] == y[i]
        if i > n // 2:
            assert abs(_wrap(step + int(st2[7, 0]), 32)) <= 1
            assert abs(_wrap(int(x[i]) + int(yi[0]), 32)) <= 4
    assert np.array_equal(st, st2)


def test_converge_narrow(oracle):
    """src/pll.rs:134-149"""
    ba = oracle.pll_from_bandwidth(8e-5, 4.0)
    n, step = 1 << 18, 0x1401235
    x = _ramp(step, n)
    st = np.zeros((9, 1), np.int32)
    half = n // 2 + 1
    oracle.pll_lanes(ba, st, x[:half], 1)
    for a in range(half, n, 4096):  # frequency checked every 4096 steps, phase on every step
        y = oracle.pll_lanes(ba, st, x[a:a + 4096], 1)
        assert abs(_wrap(step + int(st[7, 0]), 32)) <= 1 << 16
        e = (x[a:a + 4096].astype(np.int64) + y.astype(np.int64) + (1 << 31)) % (1 << 32) - (1 << 31)
        assert np.all(np.abs(e) <= 1 << 16)


def test_pll_vs_python_model(oracle):
    """random phases (wraps, clamp engaged) and random gains: oracle == big-int model, incl. state"""
    rng = np.random.default_rng(7)
    for trial in range(6):
        ba = [int(v) for v in (oracle.pll_from_bandwidth(10 ** rng.uniform(-4, -1.4), 4.0) if trial < 3
                               else rng.integers(-(1 << 31), 1 << 31, 3))]
        x = rng.integers(-(1 << 31), 1 << 31, 300).astype(np.int32) if trial % 2 else _ramp(int(rng.integers(1, 1 << 32)), 300)
        st = np.zeros((9, 1), np.int32)
        y = oracle.pll_lanes(np.array(ba, np.int32), st, x, 1)
        want, (x0, clamp, z0, y0, f0, f, yy) = _py_pll(ba, x)
        assert [int(v) for v in y] == want
        got = [int(v) for v in st[:, 0]]
        assert got == [x0, clamp, z0, y0, _wrap(f0 & 0xffffffff, 32), _wrap(f0 >> 32, 32), _wrap(f & 0xffffffff, 32),
                       _wrap(f >> 32, 32), yy]
