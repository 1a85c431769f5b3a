"""Pins the CIC part of the CPU oracle against the reference's own tests
in src/cic.rs (quickcheck properties re-run on seeded inputs, the modular-equivalence test with
its exact input vectors) and against an independent numpy model.  CPU only."""
import numpy as np
import pytest


def _zeros(oracle, N, M, dt=np.int64, lanes=1):
    return np.zeros((oracle.cic_state_words(N, M), lanes), dt)


def _modular_dec(x, N, R, M):
    """src/cic.rs:317-323: N integrators -> Downsample(R-1) (first of every R) -> N combs of delay M,
    written with numpy cumsum / shifted differences in wrapping i64 arithmetic"""
    v = np.asarray(x, np.int64).astype(np.uint64)
    for _ in range(N):
        v = np.cumsum(v, dtype=np.uint64)
    v = v[::R]
    for _ in range(N):
        d = np.concatenate([np.zeros(M, np.uint64), v[:-M] if M <= len(v) else v[:0]])[: len(v)]
        v = v - d
    return v.astype(np.int64)


def _modular_int(x, N, R, M):
    """src/cic.rs:333-339: N combs -> Hold (repeat R times) -> N integrators"""
    v = np.asarray(x, np.int64).astype(np.uint64)
    for _ in range(N):
        d = np.concatenate([np.zeros(M, np.uint64), v[:-M] if M <= len(v) else v[:0]])[: len(v)]
        v = v - d
    v = np.repeat(v, R)
    for _ in range(N):
        v = np.cumsum(v, dtype=np.uint64)
    return v.astype(np.int64)


def test_identity_dec(oracle):
    """src/cic.rs:215-222: rate 0 decimator is the identity"""
    x = np.random.default_rng(1).integers(-(1 << 62), 1 << 62, 200)
    st = _zeros(oracle, 3, 1)
    assert np.array_equal(oracle.cic_dec_lanes(3, 1, 0, st, x, 1), x)
    assert st[1, 0] == x[-1]  # get_decimate() == zoh


def test_identity_int(oracle):
    """src/cic.rs:224-232"""
    x = np.random.default_rng(2).integers(-(1 << 62), 1 << 62, 200) >> 3
    st = _zeros(oracle, 3, 1)
    assert np.array_equal(oracle.cic_int_lanes(3, 1, 0, st, x, 1), x)


@pytest.mark.parametrize("rate", [0, 1, 2, 3, 7, 15, 100])
def test_response_length_gain_settle(oracle, rate):
    """src/cic.rs:234-263: a constant input settles to x * gain after response_length outputs,
    monotonically in between; gain <= 2^gain_log2"""
    N, M = 3, 1
    shift = oracle.cic_gain_log2(N, M, rate)
    assert shift < 32 and oracle.cic_gain(N, M, rate) <= 1 << shift
    st = _zeros(oracle, N, M)
    R = rate + 1
    y_last = 0
    for x in np.random.default_rng(rate).integers(-(1 << 31), 1 << 31, 5):
        want = int(x) * oracle.cic_gain(N, M, rate)
        nfr = 2 * N + 2  # 2 * response_length = 2 * rate * N fast samples, rounded up to whole frames
        y = oracle.cic_int_lanes(N, M, rate, st, np.full(nfr, x, np.int64), 1)
        rl = oracle.cic_response_length(N, rate)
        for i, v in enumerate(y[: 2 * rl] if rl else y):
            v = int(v)
            if i < rl:
                if want > y_last:
                    assert y_last <= v < want
                elif want < y_last:
                    assert want <= v - 1 < y_last
                else:
                    assert v == want
            else:
                assert v == want
        y_last = int(y[-1])
        assert y_last == want


def test_unit_rate(oracle):
    """src/cic.rs:286-304: Cic::<i64,3,3>::new(0): gain_log2 == 6, gain == 27, ticks every sample"""
    assert oracle.cic_gain_log2(3, 3, 0) == 6 and oracle.cic_gain(3, 3, 0) == 27
    st = _zeros(oracle, 3, 3)
    x = np.array([5, -3, 1 << 20, 7, -9] + [0] * 100, np.int64)
    y = oracle.cic_dec_lanes(3, 3, 0, st, x, 1)
    assert np.array_equal(y, _modular_dec(x, 3, 1, 3))
    assert st[1, 0] == y[-1] and np.all(y[5 + 3 * 3:] == 0)


@pytest.mark.parametrize("N,R,M", [(3, 4, 1), (2, 2, 1), (3, 1, 3)])
def test_modular_decimator_matches_reference(oracle, N, R, M):
    """src/cic.rs:352-369 with the same input vector: Cic == integrators * downsample * combs"""
    x = np.array([v * 3 - 7 for v in range(-31, 65)], np.int64)
    x = x[: len(x) // R * R]
    st = _zeros(oracle, N, M)
    assert np.array_equal(oracle.cic_dec_lanes(N, M, R - 1, st, x, 1), _modular_dec(x, N, R, M))


@pytest.mark.parametrize("N,R,M", [(3, 4, 1), (2, 2, 1), (3, 1, 3)])
def test_modular_interpolator_matches_reference(oracle, N, R, M):
    """src/cic.rs:371-388 with the same input vector"""
    x = np.array([v * v - 3 * v + 2 for v in range(-12, 20)], np.int64)
    st = _zeros(oracle, N, M)
    assert np.array_equal(oracle.cic_int_lanes(N, M, R - 1, st, x, 1), _modular_int(x, N, R, M))


@pytest.mark.parametrize("dt", [np.int32, np.int64])
@pytest.mark.parametrize("layout", [0, 1])
def test_lanes_streaming_and_wrapping(oracle, dt, layout):
    """lanes == independent streams in both layouts, chunked calls == one call, i32 wraps mod 2^32"""
    rng = np.random.default_rng(5)
    N, M, rate, lanes, frames = 4, 2, 5, 5, 24
    R = rate + 1
    bits = 8 * np.dtype(dt).itemsize
    x = rng.integers(-(1 << (bits - 2)), 1 << (bits - 2), (frames, lanes, R)).astype(dt)
    flat = x.reshape(-1) if layout == 0 else np.ascontiguousarray(x.transpose(1, 0, 2)).reshape(-1)
    st = _zeros(oracle, N, M, dt, lanes)
    y = oracle.cic_dec_lanes(N, M, rate, st, flat, lanes, layout, nthreads=2)
    y = y.reshape(frames, lanes) if layout == 0 else y.reshape(lanes, frames).T
    for l in range(lanes):
        want = _modular_dec(x[:, l].reshape(-1).astype(np.int64), N, R, M)
        if dt == np.int32:
            want = want.astype(np.uint64).astype(np.uint32).astype(np.int32)  # low 32 bits
        assert np.array_equal(y[:, l], want)
    st2 = _zeros(oracle, N, M, dt, lanes)
    parts = []
    for a, b in ((0, 7), (7, 8), (8, 24)):
        xc = x[a:b].reshape(-1) if layout == 0 else np.ascontiguousarray(x[a:b].transpose(1, 0, 2)).reshape(-1)
        yc = oracle.cic_dec_lanes(N, M, rate, st2, xc, lanes, layout)
        parts.append(yc.reshape(b - a, lanes) if layout == 0 else yc.reshape(lanes, b - a).T)
    assert np.array_equal(np.concatenate(parts), y) and np.array_equal(st, st2)
