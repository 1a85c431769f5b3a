"""GPU parity of the CIC lanes against the CPU oracle: bit-exact wrapping
integer arithmetic, through the C ABI (idsp_cic_{dec,int}_{i32,i64})."""
import numpy as np
import pytest
import torch

from gpu_common import DEV, assert_bits_equal, layout_flat, to_dev, to_np

pytestmark = pytest.mark.gpu

from idsp_b200 import Cic, CicState, Lanes

NPT = {"i32": np.int32, "i64": np.int64}


def _rand(rng, kind, shape):
    bits = 8 * np.dtype(NPT[kind]).itemsize
    # large amplitudes on purpose: the integrators are meant to wrap (src/cic.rs:183-184)
    return rng.integers(-(1 << (bits - 2)), 1 << (bits - 2), shape).astype(NPT[kind])


def test_identity_and_settle_kats():
    """src/cic.rs:215-232 (rate 0 = identity, both directions) and :234-263 (settles to x * gain)"""
    rng = np.random.default_rng(1)
    x = rng.integers(-(1 << 60), 1 << 60, 200)
    for proc in (Cic(3, 1, 0).decimate(), Cic(3, 1, 0).interpolate()):
        st = CicState.default(3, 1, "i64", 1, DEV)
        y = torch.empty(200, dtype=torch.int64, device=DEV)
        Lanes(proc).block(st, to_dev(x), y)
        assert np.array_equal(to_np(y), x)
    c = Cic(3, 1, 7)
    st = CicState.default(3, 1, "i64", 1, DEV)
    nfr = 8
    y = torch.empty(nfr * 8, dtype=torch.int64, device=DEV)
    Lanes(c.interpolate()).block(st, to_dev(np.full(nfr, 12345, np.int64)), y)
    y = to_np(y)
    assert c.response_length() == 21 and c.gain() == 512 and c.gain_log2() == 9
    assert np.all(np.diff(y[:21]) >= 0) and np.all(y[21:] == 12345 * 512)


@pytest.mark.parametrize("kind", ["i32", "i64"])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("N,M,rate", [(3, 1, 15), (1, 1, 0), (2, 2, 1), (3, 3, 0), (4, 2, 6), (6, 3, 3), (5, 1, 31)])
def test_cic_dec_vs_oracle(oracle, kind, layout, N, M, rate):
    rng = np.random.default_rng(100 * N + 10 * M + rate)
    lanes, R = 131, rate + 1
    chunks = [5, 1, 17]
    x = _rand(rng, kind, (sum(chunks), lanes, R))
    so = np.zeros((oracle.cic_state_words(N, M), lanes), NPT[kind])
    st = CicState.default(N, M, kind, lanes, DEV)
    proc = Lanes(Cic(N, M, rate).decimate())
    a = 0
    for c in chunks:  # state carried across calls == streaming
        xc = layout_flat(x[a:a + c], layout)
        want = oracle.cic_dec_lanes(N, M, rate, so, xc, lanes, layout)
        y = torch.empty(c * lanes, dtype=getattr(torch, NPT[kind].__name__), device=DEV)
        proc.block(st, to_dev(xc), y, layout)
        assert_bits_equal(to_np(y), want, f"dec {kind} N={N} M={M} rate={rate} layout={layout}")
        a += c
    assert_bits_equal(st.numpy(), so, "state")


@pytest.mark.parametrize("kind", ["i32", "i64"])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("N,M,rate", [(3, 1, 15), (1, 1, 0), (2, 2, 1), (3, 3, 0), (4, 2, 6), (6, 3, 3), (5, 1, 31)])
def test_cic_int_vs_oracle(oracle, kind, layout, N, M, rate):
    rng = np.random.default_rng(200 * N + 10 * M + rate)
    lanes, R = 77, rate + 1
    chunks = [3, 1, 20]
    x = _rand(rng, kind, (sum(chunks), lanes))
    so = np.zeros((oracle.cic_state_words(N, M), lanes), NPT[kind])
    st = CicState.default(N, M, kind, lanes, DEV)
    proc = Lanes(Cic(N, M, rate).interpolate())
    a = 0
    for c in chunks:
        xc = layout_flat(x[a:a + c], layout)
        want = oracle.cic_int_lanes(N, M, rate, so, xc, lanes, layout)
        y = torch.empty(c * lanes * R, dtype=getattr(torch, NPT[kind].__name__), device=DEV)
        proc.block(st, to_dev(xc), y, layout)
        assert_bits_equal(to_np(y), want, f"int {kind} N={N} M={M} rate={rate} layout={layout}")
        a += c
    assert_bits_equal(st.numpy(), so, "state")


def test_cic_large_roundtrip_property():
    """size-independent property at a bench-like size: x16 interpolation of a constant followed by
    /16 decimation settles to x * gain_int * gain_dec on every lane"""
    lanes, frames, N, rate = 4096, 64, 3, 15
    c = Cic(N, 1, rate)
    xi = torch.full((frames * lanes,), 3, dtype=torch.int64, device=DEV)
    up = torch.empty(frames * lanes * 16, dtype=torch.int64, device=DEV)
    Lanes(c.interpolate()).block(CicState.default(N, 1, "i64", lanes, DEV), xi, up)
    dn = torch.empty(frames * lanes, dtype=torch.int64, device=DEV)
    Lanes(c.decimate()).block(CicState.default(N, 1, "i64", lanes, DEV), up, dn)
    tail = to_np(dn).reshape(frames, lanes)[8:]
    assert np.all(tail == 3 * c.gain() * c.gain())


def test_cic_argument_errors():
    from idsp_b200._lib import IdspError
    with pytest.raises(ValueError):
        Cic(7, 1, 3)
    st = CicState.default(3, 1, "i64", 4, DEV)
    with pytest.raises(ValueError):  # not a whole number of frames
        Lanes(Cic(3, 1, 3).decimate()).block(st, torch.zeros(4 * 4 + 1, dtype=torch.int64, device=DEV),
                                             torch.zeros(4, dtype=torch.int64, device=DEV))
