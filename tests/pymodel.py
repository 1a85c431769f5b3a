"""Independent pure-Python big-integer model of the integer hot-path arithmetic.

Used to cross-check the C oracle on the paths the reference itself never tests
with values (Lowpass, Lockin, DirectForm1Wide, i8/i16/i64 biquads -- SURVEY.md
section 8c "parity unpinned").  Written from the reference source with Python's
unbounded ints and explicit wrapping, i.e. a second restatement that shares no
code with oracle/idsp_oracle.c.
"""
import math


def wrap(v, bits):
    """two's complement wrap to a signed `bits`-bit integer"""
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def asr(v, n):
    return v >> n  # Python >> on ints is arithmetic


# ---- Biquad DF1 on Q<T,A,F> (src/iir/biquad.rs:366-383) -------------------
def df1_fixed(ba, F, bits, st, xs, clamp=None):
    """st = [x0,x1,y0,y1] (mutated); returns outputs."""
    A = 2 * bits
    out = []
    for x0 in xs:
        acc = wrap(ba[0] * x0 + ba[1] * st[0] + ba[2] * st[1] + ba[3] * st[2] + ba[4] * st[3], A)
        q = asr(acc, F) if F >= 0 else wrap(acc << -F, A)
        y0 = wrap(q, bits)
        st[1], st[0] = st[0], x0
        st[3], st[2] = st[2], y0
        if clamp is not None:
            u, lo, hi = clamp
            v = wrap(y0 + u, bits)
            v = lo if v < lo else (hi if v > hi else v)
            st[2] = v
            y0 = v
        out.append(y0)
    return out


# ---- DirectForm1Wide (src/iir/biquad.rs:445-480) --------------------------
def df1_wide(ba, F, sx, sy, xs, clamp=None):
    """sx = [x0,x1] i32, sy = [y0,y1] i64 (mutated)."""
    out = []
    for x0 in xs:
        acc = wrap(ba[0] * x0 + ba[1] * sx[0] + ba[2] * sx[1], 64)
        sx[1], sx[0] = sx[0], x0
        for yi, a in ((sy[0], ba[3]), (sy[1], ba[4])):
            lo = yi & 0xFFFFFFFF
            hi = wrap(yi >> 32, 32)
            acc = wrap(acc + asr(lo * a, 32), 64)
            acc = wrap(acc + hi * a, 64)
        acc = wrap(acc << (32 - F), 64)
        sy[1], sy[0] = sy[0], acc
        y0 = wrap(acc >> 32, 32)
        if clamp is not None:
            u, lo_, hi_ = clamp
            v = wrap(y0 + u, 32)
            v = lo_ if v < lo_ else (hi_ if v > hi_ else v)
            sy[0] = wrap((v << 32) | (sy[0] & 0xFFFFFFFF), 64)
            y0 = v
        out.append(y0)
    return out


# ---- DirectForm1Dither (src/iir/biquad.rs:484-538) ------------------------
def df1_dither(ba, F, st, xs, clamp=None):
    """st = [x0,x1,y0,y1,e] with e unsigned 32 bit (mutated)."""
    out = []
    for x0 in xs:
        acc = wrap(st[4] + ba[0] * x0 + ba[1] * st[0] + ba[2] * st[1] + ba[3] * st[2] + ba[4] * st[3], 64)
        acc = wrap(acc << (32 - F), 64)
        st[4] = ((acc & 0xFFFFFFFF) >> (32 - F)) if F > 0 else 0
        y0 = wrap(acc >> 32, 32)
        st[1], st[0] = st[0], x0
        st[3], st[2] = st[2], y0
        if clamp is not None:
            u, lo, hi = clamp
            v = wrap(y0 + u, 32)
            v = lo if v < lo else (hi if v > hi else v)
            st[2] = v
            y0 = v
        out.append(y0)
    return out


# ---- Lowpass<N> (src/lowpass.rs:47-78) -------------------------------------
def sat32(v):
    return max(-(1 << 31), min((1 << 31) - 1, v))


def lowpass(k, st, xs):
    out = []
    n = len(k)
    for x in xs:
        d = wrap(sat32(x - wrap(st[0] >> 32, 32)) * k[0], 64)
        if n == 1:
            st[0] = wrap(st[0] + d, 64)
            y = wrap(st[0] >> 32, 32)
            st[0] = wrap(st[0] + d, 64)
        else:
            d = wrap(d + (st[1] >> 32) * k[1], 64)
            st[1] = wrap(st[1] + d, 64)
            st[0] = wrap(st[0] + st[1], 64)
            y = wrap(st[0] >> 32, 32)
            st[0] = wrap(st[0] + st[1], 64)
            st[1] = wrap(st[1] + d, 64)
        out.append(y)
    return out


# ---- cossin (src/cossin.rs:14-67, build.rs:9-44) ---------------------------
def _rround(x):
    a = abs(x)
    r = math.floor(a)
    if a - r >= 0.5:
        r += 1
    return r if x >= 0 else -r


_COSSIN = None


def cossin_table():
    global _COSSIN
    if _COSSIN is None:
        t = []
        for i in range(128):
            a = math.pi / 4.0 * ((i + 0.5) / 128.0)
            c = _rround((math.cos(a) * 2.0 - 1.0) * 65535.0 - 1.0)
            s = _rround(math.sin(a) * 65535.0)
            t.append((c + (s << 16)) & 0xFFFFFFFF)
        _COSSIN = t
    return _COSSIN


def cossin(phase):
    lut = cossin_table()
    octant = phase & 0xFFFFFFFF
    if octant & (1 << 29):
        phase = ~phase
    phase = ((phase & 0xFFFFFFFF) << 3 & 0xFFFFFFFF) >> 10
    lookup = lut[phase >> 15]
    phase &= (1 << 15) - 1
    phase -= 1 << 14
    dphi = (phase * 51471) >> 16
    c = (lookup & 0xFFFF) + (1 << 16)
    s = lookup >> 16
    dcos = (s * dphi) >> 7
    dsin = (c * dphi) >> 8
    c = (c << 14) - dcos
    s = (s << 15) + dsin
    octant ^= octant >> 1
    if octant & (1 << 29):
        c, s = s, c
    if octant & (1 << 30):
        c = -c
    if octant & (1 << 31):
        s = -s
    return wrap(c, 32), wrap(s, 32)


# ---- Lockin (src/lockin.rs:17-39) fed by Accu (src/accu.rs:34-37) -----------
def lockin(k, accu_state, accu_step, st_i, st_q, xs):
    """returns (list of (re, im), new accu state)"""
    out = []
    ph = accu_state
    for x in xs:
        ph = wrap(ph + accu_step, 32)
        c, s = cossin(ph)
        mi = wrap((c * x) >> 32, 32)
        mq = wrap((s * x) >> 32, 32)
        (yi,) = lowpass(k, st_i, [mi])
        (yq,) = lowpass(k, st_q, [mq])
        out.append((yi, yq))
    return out, ph


# ---- atan2 (src/atan2.rs:7-82, build.rs:46-69) ------------------------------
def _divi_table():
    q31 = float(1 << 31)
    t = []
    for i in range(16):
        x0 = 1.0 + i / 16.0
        x1 = 1.0 + (i + 1) / 16.0
        t.append((_rround(q31 / x0) & 0xFFFFFFFF, _rround((1.0 / x1 - 1.0 / x0) * q31)))
    return t


_DIVI = _divi_table()


def _mul_q31(x, y):
    return ((x * y) >> 31) & 0xFFFFFFFF


def _divi(y, x):
    if x == 0:
        return 0
    shift = 32 - x.bit_length()
    y = (y << shift) & 0xFFFFFFFF
    x = (x << shift) & 0xFFFFFFFF
    fb = 27
    rem = x & ((1 << fb) - 1)
    idx = ((x << 1) & 0xFFFFFFFF) >> (1 + fb)
    base, slope = _DIVI[idx]
    step = ((slope * rem) >> fb) & 0xFFFFFFFF
    r0 = (base + step) & 0xFFFFFFFF
    return _mul_q31(y, _mul_q31(r0, (-_mul_q31(x, r0)) & 0xFFFFFFFF))


def _atani(x):
    A = [0x0517C2CD, -0x06C6496B, 0x0FBDB021, -0x25B32E0A, 0x43B34C81, -0x3BC823DD]
    x2 = wrap((x * x) >> 32, 32)
    r = 0
    for a in reversed(A):
        r = wrap((r * x2) >> 32, 32)
        r = wrap(r + a, 32)
    return ((r * x) >> 28) & 0xFFFFFFFF


def atan2(y, x):
    k = 0
    if y < 0:
        y = (1 << 31) - 1 if y == -(1 << 31) else -y
        k ^= 0xFFFFFFFF
    if x < 0:
        x = (1 << 31) - 1 if x == -(1 << 31) else -x
        k ^= 0xFFFFFFFF >> 1
    if y > x:
        y, x = x, y
        k ^= 0xFFFFFFFF >> 2
    r = _atani(_divi(y, x))
    return wrap(r ^ k, 32)
