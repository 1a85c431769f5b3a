"""GPU parity of cossin / atan2 / Lowpass / Lockin against the CPU oracle (bit-exact)."""
import numpy as np
import pytest
import torch

from gpu_common import DEV, assert_bits_equal, layout_flat, to_dev, to_np

pytestmark = pytest.mark.gpu

import idsp_b200 as ib
from idsp_b200 import Accu, Lanes, Lockin, LockinState, Lowpass, LowpassState


def test_cossin_all_2p20_phases(oracle):
    """same sweep as src/cossin.rs:130-196, compared bit-for-bit with the oracle"""
    ph = (np.arange(1 << 20, dtype=np.int64) << 12).astype(np.uint32).view(np.int32)
    got = to_np(ib.cossin(to_dev(ph)))
    assert_bits_equal(got, oracle.cossin(ph))
    assert tuple(got[0]) == (2147454703, -1898)


def test_cossin_random_and_ragged(oracle):
    rng = np.random.default_rng(1)
    for n in (1, 3, 4, 5, 1023, 100001):
        ph = rng.integers(-(1 << 31), 1 << 31, n).astype(np.int32)
        assert_bits_equal(to_np(ib.cossin(to_dev(ph))), oracle.cossin(ph))
    ph = rng.integers(-(1 << 31), 1 << 31, 5000).astype(np.int32)
    assert_bits_equal(ib.cossin(ph), oracle.cossin(ph))  # host path
    # misaligned device pointer -> scalar path
    t = to_dev(rng.integers(-(1 << 31), 1 << 31, 1001).astype(np.int32))
    assert_bits_equal(to_np(ib.cossin(t[1:])), oracle.cossin(to_np(t[1:])))


def test_atan2_exact_values_and_random(oracle):
    """src/atan2.rs:179-185 + random bit parity"""
    MAX = (1 << 31) - 1
    MIN = -(1 << 31)
    pts = np.array([[1, 0], [MAX, 0], [0, 1], [0, MAX], [0, 0], [MIN, MIN], [5, MIN], [MIN, 7], [MAX, MAX], [-1, -1]], np.int32)
    got = to_np(ib.atan2(to_dev(pts)))
    assert got[0] == 0 and got[1] == 0 and got[2] == 0x3FFFFFFF and got[3] == 0x3FFFFFFF and got[4] == 0
    assert_bits_equal(got, oracle.atan2(pts))
    rng = np.random.default_rng(2)
    for n in (1, 2, 3, 100003):
        xy = rng.integers(MIN, MAX + 1, (n, 2)).astype(np.int32)
        assert_bits_equal(to_np(ib.atan2(to_dev(xy))), oracle.atan2(xy))
    small = rng.integers(-100, 100, (5000, 2)).astype(np.int32)
    assert_bits_equal(ib.atan2(small), oracle.atan2(small))  # host path


@pytest.mark.parametrize("k", [[67465188], [1048576, -94906265], [(1 << 31) - 1], [1 << 16, -(1 << 30)]])
@pytest.mark.parametrize("layout", [0, 1])
def test_lowpass_vs_oracle(oracle, k, layout):
    rng = np.random.default_rng(3)
    for frames, lanes in [(100, 33), (64, 128), (1, 1)]:
        x = rng.integers(-(1 << 31), 1 << 31, frames * lanes).astype(np.int32)
        st0 = rng.integers(-(1 << 62), 1 << 62, (len(k), lanes)).astype(np.int64)
        so = st0.copy()
        want = oracle.lowpass_lanes(k, so, x, lanes, layout)
        st = LowpassState(to_dev(st0))
        y = torch.empty_like(to_dev(x))
        Lanes(Lowpass(k)).block(st, to_dev(x), y, layout)
        assert_bits_equal(to_np(y), want)
        assert_bits_equal(st.numpy(), so)


def test_lowpass_model_vectors():
    """SURVEY.md 8c consistency vectors (model-derived, not reference-pinned)"""
    x = torch.full((6000,), 1 << 28, dtype=torch.int32, device=DEV)
    y = torch.empty_like(x)
    Lowpass([67465188]).block(LowpassState.default(1, 1, DEV), x, y)
    yy = to_np(y)
    assert yy[:5].tolist() == [4216574, 12517255, 20557162, 28344488, 35887168] and yy[-1] == 268435456


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("layout", [0, 1])
def test_lockin_vs_oracle(oracle, order, layout):
    rng = np.random.default_rng(4 + order)
    k = [67465188] if order == 1 else [1048576, -94906265]
    for frames, lanes in [(50, 70), (128, 128), (3, 1), (52, 37), (16, 64)]:  # 52x37, 16x64: lane-major TMA tiles (frames % 4 == 0) with ragged tile / lane edges
        x = rng.integers(-(1 << 30), 1 << 30, frames * lanes).astype(np.int32)
        a0 = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
        step = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
        ao = a0.copy()
        so = np.zeros((2 * order, lanes), np.int64)
        want = oracle.lockin_lanes(k, ao, step, so, x, lanes, layout)
        st = LockinState.default(order, lanes, DEV)
        acc = Accu(to_dev(a0), to_dev(step))
        iq = torch.empty(2 * x.size, dtype=torch.int32, device=DEV)
        Lockin(Lowpass(k)).block(st, acc, to_dev(x), iq, layout)
        assert_bits_equal(to_np(iq), want)
        assert_bits_equal(to_np(acc.state), ao)
        assert_bits_equal(st.numpy(), so)
        # host path (idsp_lockin_i32_host)
        ah = a0.copy()
        sh = LockinState.default(order, lanes, None)
        iqh = np.empty(2 * x.size, np.int32)
        Lockin(Lowpass(k)).block(sh, Accu(ah, step), x, iqh, layout)
        assert_bits_equal(iqh, want)
        assert_bits_equal(ah, ao)
        assert_bits_equal(sh.words, so)


@pytest.mark.parametrize("lanes", [128, 19200 + 36])
def test_lockin_tma_kernels(oracle, lanes):
    """frame-major lock-in through the TMA kernel (4-byte in, 8-byte out boxes), narrow and wide"""
    rng = np.random.default_rng(lanes)
    k = [1048576, -94906265]
    frames = 43
    x = rng.integers(-(1 << 30), 1 << 30, frames * lanes).astype(np.int32)
    a0 = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
    step = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
    ao = a0.copy()
    so = np.zeros((4, lanes), np.int64)
    want = oracle.lockin_lanes(k, ao, step, so, x, lanes, 0, nthreads=4)
    ctx = ib.default_context(0)
    for policy in (0, 1):
        ctx.set_kernel_policy(policy)
        try:
            st = LockinState.default(2, lanes, DEV)
            acc = Accu(to_dev(a0), to_dev(step))
            iq = torch.empty(2 * x.size, dtype=torch.int32, device=DEV)
            Lockin(Lowpass(k)).block(st, acc, to_dev(x), iq, 0)
            assert_bits_equal(to_np(iq), want, f"policy={policy}")
            assert_bits_equal(to_np(acc.state), ao)
            assert_bits_equal(st.numpy(), so)
        finally:
            ctx.set_kernel_policy(0)


# ------------------------------------------------------------------ Lockin on caller-supplied phase / LO streams
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("layout", [0, 1])
def test_lockin_phase_stream_vs_oracle(oracle, order, layout):
    """`SplitProcess<(i32, Wrapping<i32>), Complex<i32>, [S; 2]>` src/lockin.rs:30-39: the phase of every sample
    comes from the caller.  Shapes cover the generic kernels and both TMA layouts (8-byte in, 8-byte out)."""
    rng = np.random.default_rng(40 + order)
    k = [67465188] if order == 1 else [1048576, -94906265]
    ctx = ib.default_context(0)
    for frames, lanes in [(50, 70), (128, 128), (3, 1), (52, 37), (16, 64), (43, 260), (64, 1024)]:
        xp = rng.integers(-(1 << 31), 1 << 31, 2 * frames * lanes).astype(np.int32)
        xp[0::2] >>= 1  # samples within +-2^30, phases over the full circle
        so = np.zeros((2 * order, lanes), np.int64)
        want = oracle.lockin_phase_lanes(k, so, xp, lanes, layout)
        for policy in (0, 1):
            ctx.set_kernel_policy(policy)
            try:
                st = LockinState.default(order, lanes, DEV)
                iq = torch.empty(xp.size, dtype=torch.int32, device=DEV)
                Lockin(Lowpass(k)).block_phase(st, to_dev(xp), iq, layout)
                assert_bits_equal(to_np(iq), want, f"policy={policy} {frames}x{lanes} ({ctx.last_kernel})")
                assert_bits_equal(st.numpy(), so)
            finally:
                ctx.set_kernel_policy(0)


def test_lockin_phase_equals_accu_form(oracle):
    """feeding the phases an Accu would produce must reproduce idsp_lockin_i32 (src/accu.rs:34-37 + lockin.rs:30-39)"""
    rng = np.random.default_rng(77)
    frames, lanes, k = 64, 96, [1048576, -94906265]
    x = rng.integers(-(1 << 30), 1 << 30, (frames, lanes)).astype(np.int32)
    a0 = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
    step = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
    ph = (a0.astype(np.int64)[None, :] + (np.arange(1, frames + 1, dtype=np.int64)[:, None] * step.astype(np.int64)[None, :]))
    ph = (ph & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
    xp = np.stack([x, ph], -1).reshape(-1)
    st1, st2 = LockinState.default(2, lanes, DEV), LockinState.default(2, lanes, DEV)
    iq1 = torch.empty(2 * x.size, dtype=torch.int32, device=DEV)
    iq2 = torch.empty_like(iq1)
    Lockin(Lowpass(k)).block(st1, Accu(to_dev(a0), to_dev(step)), to_dev(x.reshape(-1)), iq1, 0)
    Lockin(Lowpass(k)).block_phase(st2, to_dev(xp), iq2, 0)
    assert_bits_equal(to_np(iq1), to_np(iq2))
    assert_bits_equal(st1.numpy(), st2.numpy())


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("layout", [0, 1])
def test_lockin_lo_stream_vs_oracle(oracle, order, layout):
    """`SplitProcess<(X, Complex<U>), Complex<X>, [S; 2]>` src/lockin.rs:17-28 with X = i32, U = Q32<32>"""
    rng = np.random.default_rng(50 + order)
    k = [67465188] if order == 1 else [1048576, -94906265]
    for frames, lanes in [(50, 70), (3, 1), (128, 96), (40, 33)]:
        xlo = rng.integers(-(1 << 31), 1 << 31, 3 * frames * lanes).astype(np.int32)
        xlo[:9] = [np.iinfo(np.int32).min, np.iinfo(np.int32).min, np.iinfo(np.int32).max, -1, -1, 1, 0, 5, -5]
        so = np.zeros((2 * order, lanes), np.int64)
        want = oracle.lockin_lo_lanes(k, so, xlo, lanes, layout)
        st = LockinState.default(order, lanes, DEV)
        iq = torch.empty(2 * frames * lanes, dtype=torch.int32, device=DEV)
        Lockin(Lowpass(k)).block_lo(st, to_dev(xlo), iq, layout)
        assert_bits_equal(to_np(iq), want, f"{frames}x{lanes}")
        assert_bits_equal(st.numpy(), so)


def test_lockin_lo_equals_phase_form():
    """the phase impl is defined as cossin(phase) fed to the LO impl (lockin.rs:34-38)"""
    rng = np.random.default_rng(5)
    n, k = 4096, [1048576, -94906265]
    x = rng.integers(-(1 << 30), 1 << 30, n).astype(np.int32)
    ph = rng.integers(-(1 << 31), 1 << 31, n).astype(np.int32)
    cs = to_np(ib.cossin(to_dev(ph))).reshape(n, 2)
    xlo = np.stack([x, cs[:, 0], cs[:, 1]], -1).reshape(-1)
    xp = np.stack([x, ph], -1).reshape(-1)
    s1, s2 = LockinState.default(2, 1, DEV), LockinState.default(2, 1, DEV)
    a, b = torch.empty(2 * n, dtype=torch.int32, device=DEV), torch.empty(2 * n, dtype=torch.int32, device=DEV)
    Lockin(Lowpass(k)).block_lo(s1, to_dev(xlo), a, 0)
    Lockin(Lowpass(k)).block_phase(s2, to_dev(xp), b, 0)
    assert_bits_equal(to_np(a), to_np(b))


def test_cossin_exhaustive_over_every_distinct_phase(oracle):
    """cossin() ignores the low 7 phase bits ((phase << 3) >> 10, src/cossin.rs:22-25): 2^25 distinct inputs.  All of
    them (random low bits) through the full-circle table kernel against the oracle -- the table folds the octant
    reversal, swap and negations of cossin.rs:17-21, 56-65 into 1024 pre-mapped entries, so every octant boundary
    and every table edge is covered here."""
    rng = np.random.default_rng(25)
    for part in range(4):
        hi = np.arange(part << 23, (part + 1) << 23, dtype=np.int64)
        ph = ((hi << 7) | rng.integers(0, 128, hi.size)).astype(np.uint32).view(np.int32)
        got = to_np(ib.cossin(to_dev(ph)))
        assert_bits_equal(got, oracle.cossin(ph), f"part {part}")


# ------------------------------------------------------------------ speculative saturation of the tile kernels
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("layout", [0, 1])
def test_lockin_saturating_states_take_the_exact_path(oracle, order, layout):
    """The TMA lock-in kernels subtract without saturation while the low-pass state's high word is inside
    [-2^30, 2^30) and redo a tile with `saturating_sub` (src/lowpass.rs:56) otherwise (ops.cuh,
    IDSP_LOCKIN_SPEC_MEMBERS).  Start a third of the lanes from states at / near full scale -- so that the
    subtraction really saturates -- drive full-scale samples, and compare every output and the final state
    with the oracle; the other lanes stay on the speculative path in the same warps."""
    rng = np.random.default_rng(900 + 2 * order + layout)
    k = [67465188] if order == 1 else [1048576, -94906265]
    ctx = ib.default_context(0)
    for frames, lanes in [(64, 256), (48, 132), (16, 64)]:
        x = rng.integers(-(1 << 31), 1 << 31, frames * lanes).astype(np.int32)
        a0 = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
        step = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
        s0 = np.zeros((2 * order, lanes), np.int64)
        hot = rng.random(lanes) < 0.34
        big = rng.integers(-(1 << 63), (1 << 63) - 1, (2 * order, lanes), dtype=np.int64)
        edge = np.array([np.iinfo(np.int64).min, np.iinfo(np.int64).max, (1 << 62), -(1 << 62), (1 << 62) - 1, -(1 << 62) - 1], np.int64)
        big[:, : min(lanes, edge.size)] = edge[: min(lanes, edge.size)]
        hot[: min(lanes, edge.size)] = True
        s0[:, hot] = big[:, hot]
        # Accu form
        ao, so = a0.copy(), s0.copy()
        want = oracle.lockin_lanes(k, ao, step, so, x, lanes, layout)
        st = LockinState.default(order, lanes, DEV)
        st.words.copy_(torch.from_numpy(s0).to(DEV))
        acc = Accu(to_dev(a0), to_dev(step))
        iq = torch.empty(2 * x.size, dtype=torch.int32, device=DEV)
        Lockin(Lowpass(k)).block(st, acc, to_dev(x), iq, layout)
        assert "tma" in ctx.last_kernel, ctx.last_kernel
        assert_bits_equal(to_np(iq), want, f"accu form {frames}x{lanes}")
        assert_bits_equal(st.numpy(), so)
        assert_bits_equal(to_np(acc.state), ao)
        # (x, phase) form
        xp = np.empty(2 * x.size, np.int32)
        xp[0::2] = x
        xp[1::2] = rng.integers(-(1 << 31), 1 << 31, x.size).astype(np.int32)
        so = s0.copy()
        want = oracle.lockin_phase_lanes(k, so, xp, lanes, layout)
        st = LockinState.default(order, lanes, DEV)
        st.words.copy_(torch.from_numpy(s0).to(DEV))
        iq = torch.empty(xp.size, dtype=torch.int32, device=DEV)
        Lockin(Lowpass(k)).block_phase(st, to_dev(xp), iq, layout)
        assert "tma" in ctx.last_kernel, ctx.last_kernel
        assert_bits_equal(to_np(iq), want, f"phase form {frames}x{lanes}")
        assert_bits_equal(st.numpy(), so)
