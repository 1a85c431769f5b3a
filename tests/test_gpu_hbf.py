"""GPU parity of the half-band FIR family against the CPU oracle (bit-exact f32, no FMA)."""
import numpy as np
import pytest
import torch

from gpu_common import DEV, assert_bits_equal, layout_flat, to_dev, to_np

pytestmark = pytest.mark.gpu

import idsp_b200 as ib
from idsp_b200 import (EvenAntiSymmetric, EvenSymmetric, HbfDec, HbfDecCascade, HbfInt, HbfIntCascade,
                       Lanes, OddAntiSymmetric, OddSymmetric, Split, hbf_taps)
from idsp_b200.hbf import (FirState, HbfDec16, HbfInt16, hbf_dec_response_length,
                           hbf_int_response_length, _dec_state, _int_state)


def test_kat_hbf_dec_simple():
    """src/hbf.rs:548-556"""
    h = Split.new(EvenSymmetric([0.5]), HbfDec.default(1, 1, DEV))
    h.block(torch.empty(0, dtype=torch.float32, device=DEV), torch.empty(0, dtype=torch.float32, device=DEV))
    y = torch.zeros(4, dtype=torch.float32, device=DEV)
    h.block(to_dev(np.ones(8, np.float32)), y)
    assert to_np(y).tolist() == [1.5, 2.0, 2.0, 2.0]


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("layout", [0, 1])
def test_single_stage_dec_int(oracle, idx, layout):
    rng = np.random.default_rng(idx)
    taps = hbf_taps()[idx]
    M = len(taps)
    for n, lanes in [(1, 1), (37, 5), (64, 33), (100, 128)]:
        x = rng.standard_normal((n, lanes, 2)).astype(np.float32)
        st0 = rng.standard_normal((3 * M - 2, lanes)).astype(np.float32)
        so = st0.copy()
        want = np.empty((n, lanes), np.float32)
        for l in range(lanes):
            s = so[:, l].copy()
            want[:, l] = oracle.hbf_dec(taps, s, x[:, l].reshape(-1))
            so[:, l] = s
        st = HbfDec(to_dev(st0), M)
        y = torch.empty(n * lanes, dtype=torch.float32, device=DEV)
        Lanes(EvenSymmetric(taps)).inner.block(st, to_dev(layout_flat(x, layout)), y, layout)
        assert_bits_equal(to_np(y), layout_flat(want, layout), f"dec idx={idx}")
        assert_bits_equal(st.numpy(), so)
        # interpolator
        xi = rng.standard_normal((n, lanes)).astype(np.float32)
        si0 = rng.standard_normal((2 * M - 1, lanes)).astype(np.float32)
        sio = si0.copy()
        wanti = np.empty((n, lanes, 2), np.float32)
        for l in range(lanes):
            s = sio[:, l].copy()
            wanti[:, l] = oracle.hbf_int(taps, s, xi[:, l].copy()).reshape(n, 2)
            sio[:, l] = s
        sti = HbfInt(to_dev(si0), M)
        yi = torch.empty(2 * n * lanes, dtype=torch.float32, device=DEV)
        EvenSymmetric(taps).block(sti, to_dev(layout_flat(xi, layout)), yi, layout)
        assert_bits_equal(to_np(yi), layout_flat(wanti, layout), f"int idx={idx}")
        assert_bits_equal(sti.numpy(), sio)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("layout", [0, 1])
def test_dec_cascade_vs_oracle(oracle, k, layout):
    rng = np.random.default_rng(20 + k)
    R = 1 << k
    W = oracle.hbf_dec_state_words(k)
    for n_out, lanes in [(1, 1), (9, 3), (40, 70), (64, 128), (130, 257), (256, 64), (512, 32)]:
        x = rng.uniform(-1, 1, (n_out, lanes, R)).astype(np.float32)
        st0 = rng.uniform(-1, 1, (W, lanes)).astype(np.float32)
        so = st0.copy()
        want = oracle.hbf_dec_cascade_lanes(k, so, layout_flat(x, layout), lanes, layout)
        st = _dec_state(k)(lanes, DEV)
        st.words.copy_(to_dev(st0))
        y = torch.empty(n_out * lanes, dtype=torch.float32, device=DEV)
        Lanes(HbfDecCascade(k)).block(st, to_dev(layout_flat(x, layout)), y, layout)
        assert_bits_equal(to_np(y), want, f"dec cascade k={k} layout={layout} {n_out}x{lanes}")
        assert_bits_equal(st.numpy(), so, "state")


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("layout", [0, 1])
def test_int_cascade_vs_oracle(oracle, k, layout):
    rng = np.random.default_rng(30 + k)
    W = oracle.hbf_int_state_words(k)
    for n_in, lanes in [(1, 1), (9, 3), (40, 70), (64, 128)]:
        x = rng.uniform(-1, 1, (n_in, lanes)).astype(np.float32)
        st0 = rng.uniform(-1, 1, (W, lanes)).astype(np.float32)
        so = st0.copy()
        want = oracle.hbf_int_cascade_lanes(k, so, layout_flat(x, layout), lanes, layout)
        st = _int_state(k)(lanes, DEV)
        st.words.copy_(to_dev(st0))
        y = torch.empty(n_in * lanes << k, dtype=torch.float32, device=DEV)
        Lanes(HbfIntCascade(k)).block(st, to_dev(layout_flat(x, layout)), y, layout)
        assert_bits_equal(to_np(y), want, f"int cascade k={k}")
        assert_bits_equal(st.numpy(), so, "state")


def test_kat_response_lengths():
    """src/hbf.rs:576-609 on the GPU"""
    rng = np.random.default_rng(1)
    h = HbfDec16(1, DEV)
    c = HbfDecCascade(4)
    y = torch.empty(100, dtype=torch.float32, device=DEV)
    c.block(h, to_dev(rng.random(100 << 4, dtype=np.float32)), y)
    y = torch.empty(64, dtype=torch.float32, device=DEV)
    c.block(h, torch.zeros(1 << 10, dtype=torch.float32, device=DEV), y)
    n = hbf_dec_response_length(4)
    yy = to_np(y)
    assert yy[n - 1] != 0.0 and yy[n] == 0.0
    r = hbf_int_response_length(4)
    x = np.zeros((r >> 4) + 1, np.float32); x[0] = 1.0
    yi = torch.empty(x.size << 4, dtype=torch.float32, device=DEV)
    HbfIntCascade(4).block(HbfInt16(1, DEV), to_dev(x), yi)
    yi = to_np(yi)
    assert yi[r] != 0.0 and np.all(yi[r + 1:] == 0.0)


def test_dec_cascade_streaming_blocks(oracle):
    """output independent of how the stream is chunked into block() calls"""
    rng = np.random.default_rng(3)
    lanes, n_out, k = 96, 192, 4
    x = rng.uniform(-1, 1, (n_out, lanes, 16)).astype(np.float32)
    so = np.zeros((oracle.hbf_dec_state_words(k), lanes), np.float32)
    want = oracle.hbf_dec_cascade_lanes(k, so, x.reshape(-1), lanes, 0).reshape(n_out, lanes)
    st = HbfDec16(lanes, DEV)
    outs = []
    for a, b in ((0, 64), (64, 65), (65, 192)):
        y = torch.empty((b - a) * lanes, dtype=torch.float32, device=DEV)
        Lanes(HbfDecCascade(k)).block(st, to_dev(x[a:b].reshape(-1)), y)
        outs.append(to_np(y).reshape(b - a, lanes))
    assert_bits_equal(np.concatenate(outs), want)
    assert_bits_equal(st.numpy(), so)


def test_dec_cascade_host_buffers(oracle):
    rng = np.random.default_rng(4)
    for layout in (0, 1):
        lanes, n_out, k = 160, 64, 4
        x = rng.uniform(-1, 1, n_out * lanes * 16).astype(np.float32)
        so = np.zeros((oracle.hbf_dec_state_words(k), lanes), np.float32)
        want = oracle.hbf_dec_cascade_lanes(k, so, x, lanes, layout)
        st = HbfDec16(lanes, None)
        y = np.empty(n_out * lanes, np.float32)
        Lanes(HbfDecCascade(k)).block(st, x, y, layout)
        assert_bits_equal(y, want)
        assert_bits_equal(st.words, so)


@pytest.mark.parametrize("cls,odd,sym", [(OddSymmetric, 1, 1), (EvenSymmetric, 0, 1), (OddAntiSymmetric, 1, 0), (EvenAntiSymmetric, 0, 0)])
def test_single_rate_fir_types(oracle, cls, odd, sym):
    """type_fir! (src/hbf.rs:70-138)"""
    rng = np.random.default_rng(40 + odd * 2 + sym)
    taps = rng.standard_normal(6).astype(np.float32)
    M = 6
    lanes, n = 19, 70
    x = rng.standard_normal((n, lanes)).astype(np.float32)
    W = 2 * M - 1 + odd
    so = np.zeros((W, lanes), np.float32)
    want = np.empty((n, lanes), np.float32)
    for l in range(lanes):
        s = so[:, l].copy()
        want[:, l] = oracle.fir(taps, odd, sym, s, x[:, l].copy())
        so[:, l] = s
    for layout in (0, 1):
        st = FirState.default(W, lanes, DEV)
        y = torch.empty(n * lanes, dtype=torch.float32, device=DEV)
        cls(taps).block(st, to_dev(layout_flat(x, layout)), y, layout)
        assert_bits_equal(to_np(y), layout_flat(want, layout))
        assert_bits_equal(st.numpy(), so)


@pytest.mark.parametrize("k", [1, 4])
def test_chain_vs_oracle(oracle, k):
    """config 5: HbfDec -> HbfInt -> Biquad DF1 f32 fused"""
    from idsp_b200 import Biquad, Filter
    rng = np.random.default_rng(50 + k)
    ba = Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), "f32").ba
    R = 1 << k
    W = oracle.hbf_dec_state_words(k) + oracle.hbf_int_state_words(k) + 4
    ctx = ib.default_context(0)
    for layout in (0, 1):
        lanes, n_low = 50, 40
        x = rng.uniform(-1, 1, n_low * lanes * R).astype(np.float32)
        so = np.zeros((W, lanes), np.float32)
        want = oracle.chain_lanes(k, ba, so, x, lanes, layout)
        st = torch.zeros((W, lanes), dtype=torch.float32, device=DEV)
        y = ctx.chain(k, ba, st, to_dev(x), lanes=lanes, layout=layout)
        assert_bits_equal(to_np(y), want)
        assert_bits_equal(to_np(st), so)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_chain_tiled_fused_biquad_streaming(oracle, k):
    """config 5, lane-major, whole tiles: tiled decimator -> tiled interpolator with the biquad warp
    (two passes); state carried over calls of 8, 1, 9 and 11 tiles and a ragged call (the short and
    ragged ones run on the other kernels), ragged lane count, == the oracle called the same way"""
    from idsp_b200 import Biquad, Filter
    rng = np.random.default_rng(70 + k)
    ba = Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), "f32").ba
    R = 1 << k
    TI = 512 >> k
    W = oracle.hbf_dec_state_words(k) + oracle.hbf_int_state_words(k) + 4
    ctx = ib.default_context(0)
    for lanes in (50, 8, 3):
        so = np.zeros((W, lanes), np.float32)
        st = torch.zeros((W, lanes), dtype=torch.float32, device=DEV)
        for n_low in (8 * TI, TI, 9 * TI, TI + 5, 11 * TI):  # >= 8 tiles: fused path; fewer / ragged: the other kernels
            x = rng.uniform(-1, 1, n_low * lanes * R).astype(np.float32)
            want = oracle.chain_lanes(k, ba, so, x, lanes, 1)
            y = ctx.chain(k, ba, st, to_dev(x), lanes=lanes, layout=1)
            assert_bits_equal(to_np(y), want, f"k={k} lanes={lanes} n_low={n_low}")
            assert_bits_equal(to_np(st), so, f"state k={k} lanes={lanes} n_low={n_low}")


@pytest.mark.parametrize("k", [2, 4, 5])
def test_chain_tiled_fused_biquad_many_lanes(oracle, k):
    """above 8192 lanes the fused path uses 16-lane x 256-output tiles (ragged last CTA: 8200 = 512 * 16 + 8)"""
    from idsp_b200 import Biquad, Filter
    rng = np.random.default_rng(90 + k)
    ba = Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), "f32").ba
    R, lanes = 1 << k, 8200
    W = oracle.hbf_dec_state_words(k) + oracle.hbf_int_state_words(k) + 4
    ctx = ib.default_context(0)
    so = np.zeros((W, lanes), np.float32)
    st = torch.zeros((W, lanes), dtype=torch.float32, device=DEV)
    for n_low in (8 * (256 >> k), 9 * (256 >> k)):
        x = rng.uniform(-1, 1, n_low * lanes * R).astype(np.float32)
        want = oracle.chain_lanes(k, ba, so, x, lanes, 1)
        y = ctx.chain(k, ba, st, to_dev(x), lanes=lanes, layout=1)
        assert_bits_equal(to_np(y), want, f"k={k} n_low={n_low}")
        assert_bits_equal(to_np(st), so, f"state k={k} n_low={n_low}")


@pytest.mark.parametrize("layout", [1, 0])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_dec_cascade_tiled_kernel_streaming(oracle, k, layout):
    """the tiled kernel (lane-major: TMA rows; frame-major: LDGSTS gathers, k >= 2) + generic tail,
    state carried across ragged calls, == one oracle pass; also the generic-only path (policy 1) and
    the packed f32x2 variant (policy 3) give the same bits"""
    rng = np.random.default_rng(60 + k)
    R = 1 << k
    TO = 512 >> k
    lanes = 37
    chunks = [2 * TO + 3, TO, 1, 3 * TO - 1, 6 * TO + 2]
    n_out = sum(chunks)
    x = rng.uniform(-1, 1, (n_out, lanes, R)).astype(np.float32)  # [frames, lanes, R]
    so = np.zeros((oracle.hbf_dec_state_words(k), lanes), np.float32)
    want = oracle.hbf_dec_cascade_lanes(k, so, layout_flat(x, 1), lanes, 1).reshape(lanes, n_out).T
    ctx = ib.default_context(0)
    for policy in (0, 1, 3):  # default tiled kernel, generic only, packed f32x2 tiled variant
        ctx.set_kernel_policy(policy)
        try:
            st = _dec_state(k)(lanes, DEV)
            outs, a = [], 0
            for c in chunks:
                y = torch.empty(lanes * c, dtype=torch.float32, device=DEV)
                Lanes(HbfDecCascade(k)).block(st, to_dev(layout_flat(x[a:a + c], layout)), y, layout)
                outs.append(to_np(y).reshape(c, lanes) if layout == 0 else to_np(y).reshape(lanes, c).T)
                a += c
            assert_bits_equal(np.concatenate(outs, axis=0), want, f"k={k} policy={policy} layout={layout}")
            assert_bits_equal(st.numpy(), so, "state")
        finally:
            ctx.set_kernel_policy(0)


@pytest.mark.parametrize("layout", [1, 0])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_int_cascade_tiled_kernel_streaming(oracle, k, layout):
    """x2^k: tiled kernel (lane-major: TMA bulk stores; frame-major: 16-byte piece stores, k >= 2) +
    generic tail, ragged calls with carried state == one oracle pass; the generic-only path
    (policy 1) gives the same bits"""
    rng = np.random.default_rng(70 + k)
    R = 1 << k
    TI = 512 >> k
    lanes = 21
    chunks = [2 * TI + 4, TI, 4, 3 * TI - 4, 5, 4 * TI]
    n_in = sum(chunks)
    x = rng.uniform(-1, 1, (n_in, lanes)).astype(np.float32)  # [frames, lanes]
    so = np.zeros((oracle.hbf_int_state_words(k), lanes), np.float32)
    want = oracle.hbf_int_cascade_lanes(k, so, layout_flat(x, 1), lanes, 1).reshape(lanes, n_in, R).transpose(1, 0, 2)
    ctx = ib.default_context(0)
    for policy in (0, 1):
        ctx.set_kernel_policy(policy)
        try:
            st = _int_state(k)(lanes, DEV)
            outs, a = [], 0
            for c in chunks:
                y = torch.empty(lanes * c * R, dtype=torch.float32, device=DEV)
                Lanes(HbfIntCascade(k)).block(st, to_dev(layout_flat(x[a:a + c], layout)), y, layout)
                yy = to_np(y)
                outs.append(yy.reshape(c, lanes, R) if layout == 0 else yy.reshape(lanes, c, R).transpose(1, 0, 2))
                a += c
            assert_bits_equal(np.concatenate(outs, axis=0), want, f"k={k} policy={policy} layout={layout}")
            assert_bits_equal(st.numpy(), so, "state")
        finally:
            ctx.set_kernel_policy(0)


# ------------------------------------------------------------------ caller-supplied tap sets (HBF_TAPS_98, arbitrary)
def _oracle_dec_cascade_taps(oracle, taps, so, x_lane):
    """sequential composition of the reference's single stages, highest rate first (hbf.rs:385-421);
    so = this lane's state words in processing order, updated in place"""
    k = len(taps)
    off, v = 0, x_lane
    for s in range(k):
        t = taps[k - 1 - s]
        w = 3 * len(t) - 2
        st = so[off:off + w].copy()
        v = oracle.hbf_dec(t, st, v)
        so[off:off + w] = st
        off += w
    return v


def _oracle_int_cascade_taps(oracle, taps, so, x_lane):
    off, v = 0, x_lane
    for s in range(len(taps)):
        t = taps[s]
        w = 2 * len(t) - 1
        st = so[off:off + w].copy()
        v = oracle.hbf_int(t, st, v)
        so[off:off + w] = st
        off += w
    return v


def test_hbf_taps_98_values():
    """HBF_TAPS_98 transcribed from src/hbf.rs:258-292"""
    t = ib.hbf_taps_98()
    assert [len(v) for v in t] == [15, 6, 3, 3, 2]
    assert t[0][0] == np.float32(7.02144012e-05) and t[0][-1] == np.float32(6.33592923e-01)
    assert t[1].tolist() == [np.float32(v) for v in (-0.00086943, 0.00577837, -0.02201674, 0.06357869, -0.16627679, 0.61979312)]
    assert t[4].tolist() == [np.float32(-0.06291796), np.float32(0.5629161)]
    for v in t:  # half-band: DC gain of the /2 stage is 2 (no 0.5 scaling)
        assert abs(2 * float(np.sum(v.astype(np.float64))) + 1 - 2) < 2e-4


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("tapset", ["taps98", "taps98_whole_tiles", "random", "builtin"])
def test_cascade_with_caller_taps(oracle, k, layout, tapset):
    """taps98: ragged length (frame-major: tiled kernels + stage-by-stage tail, lane-major: stage by stage);
    taps98_whole_tiles: the tiled kernels compiled for HBF_TAPS_98 in both layouts; random: arbitrary run-time taps"""
    rng = np.random.default_rng(k * 7 + layout)
    whole = tapset == "taps98_whole_tiles"
    if whole:
        tapset = "taps98"
    if tapset == "taps98":
        taps = list(ib.hbf_taps_98()[:k])
    elif tapset == "builtin":  # HBF_TAPS passed explicitly must take the built-in (tiled) path and agree with it
        taps = list(hbf_taps()[:k])
    else:
        taps = [(rng.standard_normal(m) * 0.3).astype(np.float32) for m in (7, 1, 12, 32, 9)[:k]]
    R = 1 << k
    lanes, n_out = 20, (1024 >> k) * (2 if whole else 1) + (0 if whole else 3)
    xl = rng.uniform(-1, 1, (lanes, n_out * R)).astype(np.float32)
    xf = np.ascontiguousarray(xl.reshape(lanes, n_out, R).swapaxes(0, 1)).reshape(-1) if layout == 0 else xl.reshape(-1)
    cfg = HbfDecCascade(k, taps)
    st = cfg.state(lanes, DEV)
    so = np.zeros(tuple(st.words.shape), np.float32)
    want = np.stack([_oracle_dec_cascade_taps(oracle, taps, so[:, l], xl[l]) for l in range(lanes)])  # [lanes][n_out]
    y = torch.empty(n_out * lanes, dtype=torch.float32, device=DEV)
    # streamed in two calls: the state carries over like repeated block() calls
    cut = (n_out // 2)
    if layout == 0:
        Lanes(cfg).block(st, to_dev(xf[: cut * lanes * R]), y[: cut * lanes], layout)
        Lanes(cfg).block(st, to_dev(xf[cut * lanes * R:]), y[cut * lanes:], layout)
        assert_bits_equal(to_np(y), np.ascontiguousarray(want.T).reshape(-1), f"dec {tapset} k={k}")
    else:
        Lanes(cfg).block(st, to_dev(xf), y, layout)
        assert_bits_equal(to_np(y), want.reshape(-1), f"dec {tapset} k={k}")
    assert_bits_equal(st.numpy(), so, "dec state")
    # interpolator
    cfi = HbfIntCascade(k, taps)
    sti = cfi.state(lanes, DEV)
    si = np.zeros(tuple(sti.words.shape), np.float32)
    want_i = np.stack([_oracle_int_cascade_taps(oracle, taps, si[:, l], want[l]) for l in range(lanes)])  # [lanes][n_out*R]
    xin = np.ascontiguousarray(want.T).reshape(-1) if layout == 0 else want.reshape(-1)
    yi = torch.empty(n_out * lanes * R, dtype=torch.float32, device=DEV)
    Lanes(cfi).block(sti, to_dev(xin), yi, layout)
    wf = np.ascontiguousarray(want_i.reshape(lanes, n_out, R).swapaxes(0, 1)).reshape(-1) if layout == 0 else want_i.reshape(-1)
    assert_bits_equal(to_np(yi), wf, f"int {tapset} k={k}")
    assert_bits_equal(sti.numpy(), si, "int state")


def test_builtin_taps_passed_explicitly_use_tiled_kernels():
    ctx = ib.default_context(0)
    k, lanes, n_out = 4, 16, 256
    x = torch.empty(lanes * n_out * 16, dtype=torch.float32, device=DEV).uniform_(-1, 1)
    y = torch.empty(lanes * n_out, dtype=torch.float32, device=DEV)
    cfg = HbfDecCascade(k, hbf_taps()[:k])
    Lanes(cfg).block(cfg.state(lanes, DEV), x, y, 1)
    assert ctx.last_kernel.startswith("hbf tiled"), ctx.last_kernel
    cfg = HbfDecCascade(k, ib.hbf_taps_98()[:k])
    Lanes(cfg).block(cfg.state(lanes, DEV), x, y, 1)
    assert ctx.last_kernel.startswith("hbf tiled"), ctx.last_kernel          # HBF_TAPS_98 is compiled in too
    Lanes(cfg).block(cfg.state(lanes, DEV), x[: lanes * 250 * 16], y[: lanes * 250], 1)
    assert ctx.last_kernel.startswith("hbf single stage"), ctx.last_kernel   # ragged lane-major: stage by stage
    other = [t * np.float32(0.5) for t in ib.hbf_taps_98()[:k]]
    cfg = HbfDecCascade(k, other)
    Lanes(cfg).block(cfg.state(lanes, DEV), x, y, 1)
    assert ctx.last_kernel.startswith("hbf single stage"), ctx.last_kernel   # any other tap set: run-time taps


# ------------------------------------------------------------------ misaligned caller pointers
@pytest.mark.parametrize("layout", [0, 1])
def test_hbf_misaligned_pointers(oracle, layout):
    """views that are only 4-byte aligned (x[1:], y[1:]) must take the scalar branches, never fault"""
    rng = np.random.default_rng(3)
    lanes = 24
    for k in (1, 2, 4):
        R = 1 << k
        n_out = (1024 >> k) + 1
        xl = rng.uniform(-1, 1, (lanes, n_out * R)).astype(np.float32)
        xf = np.ascontiguousarray(xl.reshape(lanes, n_out, R).swapaxes(0, 1)).reshape(-1) if layout == 0 else xl.reshape(-1)
        so = np.zeros((oracle.hbf_dec_state_words(k), lanes), np.float32)
        want = oracle.hbf_dec_cascade_lanes(k, so, xf, lanes, layout)
        xb = torch.zeros(xf.size + 1, dtype=torch.float32, device=DEV)
        xb[1:] = to_dev(xf)
        yb = torch.zeros(want.size + 1, dtype=torch.float32, device=DEV)
        st = _dec_state(k)(lanes, DEV)
        Lanes(HbfDecCascade(k)).block(st, xb[1:], yb[1:], layout)
        assert_bits_equal(to_np(yb[1:]), want, f"dec /{R} misaligned")
        si = np.zeros((oracle.hbf_int_state_words(k), lanes), np.float32)
        want_i = oracle.hbf_int_cascade_lanes(k, si, want, lanes, layout)
        yi = torch.zeros(want_i.size + 1, dtype=torch.float32, device=DEV)
        Lanes(HbfIntCascade(k)).block(_int_state(k)(lanes, DEV), yb[1:], yi[1:], layout)
        assert_bits_equal(to_np(yi[1:]), want_i, f"int x{R} misaligned")
    # single stages and the chain
    taps = hbf_taps()[2]
    M = len(taps)
    x = rng.uniform(-1, 1, (lanes, 64)).astype(np.float32)
    xf = np.ascontiguousarray(x.reshape(lanes, 32, 2).swapaxes(0, 1)).reshape(-1) if layout == 0 else x.reshape(-1)
    want_l = [oracle.hbf_dec(taps, np.zeros(3 * M - 2, np.float32), x[l]) for l in range(lanes)]
    want = np.stack(want_l, 1).reshape(-1) if layout == 0 else np.concatenate(want_l)
    xb = torch.zeros(xf.size + 1, dtype=torch.float32, device=DEV)
    xb[1:] = to_dev(xf)
    yb = torch.zeros(want.size + 1, dtype=torch.float32, device=DEV)
    EvenSymmetric(taps).block(HbfDec.default(M, lanes, DEV), xb[1:], yb[1:], layout)
    assert_bits_equal(to_np(yb[1:]), want, "single stage misaligned")
    from idsp_b200 import Biquad, Filter, _lib
    W = int(_lib.lib().idsp_chain_state_words(2))
    ba = np.asarray(Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), "f32").ba)
    xl = rng.uniform(-1, 1, (lanes, 40 * 4)).astype(np.float32)
    xf = np.ascontiguousarray(xl.reshape(lanes, 40, 4).swapaxes(0, 1)).reshape(-1) if layout == 0 else xl.reshape(-1)
    so = np.zeros((W, lanes), np.float32)
    want = oracle.chain_lanes(2, ba, so, xf, lanes, layout)
    xb = torch.zeros(xf.size + 1, dtype=torch.float32, device=DEV)
    xb[1:] = to_dev(xf)
    yb = torch.zeros(xf.size + 1, dtype=torch.float32, device=DEV)
    ib.default_context(0).chain(2, ba, torch.zeros((W, lanes), dtype=torch.float32, device=DEV), xb[1:], yb[1:], lanes=lanes, layout=layout)
    assert_bits_equal(to_np(yb[1:]), want, "chain misaligned")


def test_state_validation_errors():
    """a state built for other taps / another operator is rejected on the host instead of being read out of bounds"""
    x = torch.zeros(64, dtype=torch.float32, device=DEV)
    y = torch.zeros(32, dtype=torch.float32, device=DEV)
    with pytest.raises(TypeError):
        EvenSymmetric(hbf_taps()[0]).block(HbfDec.default(3, 1, DEV), x, y)          # M = 3 state, 23 taps
    with pytest.raises(TypeError):
        OddSymmetric([0.1, 0.2]).block(FirState.default(3, 1, DEV), y, torch.zeros_like(y))  # LEN = 4 needed
    ctx = ib.default_context(0)
    with pytest.raises(ValueError):
        ctx.hbf_dec_cascade(4, torch.zeros((10, 1), dtype=torch.float32, device=DEV), x, torch.zeros(4, dtype=torch.float32, device=DEV), lanes=1)
    with pytest.raises(ValueError):
        ctx.biquad("df1", [1, 0, 0, 0, 0], 0, None, torch.zeros((4, 3), dtype=torch.float32, device=DEV), x, None, lanes=3)  # 64 % 3
    with pytest.raises(TypeError):
        ctx.biquad("df1", [1, 0, 0, 0, 0], 0, None, torch.zeros((4, 2), dtype=torch.int32, device=DEV), x, None, lanes=2)  # i32 state, f32 samples
    with pytest.raises(ValueError):
        ctx.biquad("df1", [1, 0, 0, 0, 0], 0, None, torch.zeros((4, 2), dtype=torch.float32, device=DEV), x, y, lanes=2)   # short output


def test_chain_host_buffers(oracle):
    """idsp_chain_f32_host: numpy in / out, one PCIe round trip for dec -> int -> biquad"""
    from idsp_b200 import Biquad, Filter, _lib
    rng = np.random.default_rng(9)
    k = 4
    W = int(_lib.lib().idsp_chain_state_words(k))
    ba = np.asarray(Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), "f32").ba)
    ctx = ib.default_context(0)
    for layout, lanes, n_low in ((1, 300, 64), (0, 40, 70), (1, 4096, 256)):
        x = rng.uniform(-1, 1, lanes * n_low * 16).astype(np.float32)
        so = np.zeros((W, lanes), np.float32)
        want = oracle.chain_lanes(k, ba, so, x, lanes, layout, nthreads=4)
        st = np.zeros((W, lanes), np.float32)
        y = np.empty_like(x)
        ctx.chain(k, ba, st, x, y, lanes=lanes, layout=layout)
        assert_bits_equal(y, want, f"chain host layout={layout}")
        assert_bits_equal(st, so, "chain host state")


@pytest.mark.parametrize("k", [2, 3, 4, 5])
@pytest.mark.parametrize("taps98", [False, True])
@pytest.mark.parametrize("lanes", [1, 2, 9, 64, 131])
def test_dec16_frame_major_tensor_map_input(oracle, lanes, taps98, k):
    """frame-major /16 (`[[f32; 16]; lanes]` frames, BASELINE configs[2]): the input tiles of the tiled kernel are
    tensor-map boxes of one lane pair x 33 frames in the 128-byte swizzled layout (hbf_fast_scalar_body.cuh, FmTma).
    Odd lane counts (a box half / fully outside the tensor), one lane, several CTAs; whole tiles + a ragged tail;
    state carried over three calls; both compiled tap sets; == the oracle bit for bit."""
    rng = np.random.default_rng(4242 + lanes + k)
    R, TO = 1 << k, 512 >> k
    chunks = [3 * TO, 2 * TO + 5, 4 * TO]
    n_out = sum(chunks)
    x = rng.uniform(-1, 1, (n_out, lanes, R)).astype(np.float32)
    ctx = ib.default_context(0)
    if taps98:
        taps = list(ib.hbf_taps_98()[:k])
        casc = HbfDecCascade(k, taps)
        st = casc.state(lanes, DEV)
        so = np.zeros(tuple(st.words.shape), np.float32)
        xl = np.ascontiguousarray(x.swapaxes(0, 1)).reshape(lanes, n_out * R)  # [lanes][samples]
        want = np.stack([_oracle_dec_cascade_taps(oracle, taps, so[:, l], xl[l]) for l in range(lanes)]).T
    else:
        casc = HbfDecCascade(k)
        so = np.zeros((oracle.hbf_dec_state_words(k), lanes), np.float32)
        want = oracle.hbf_dec_cascade_lanes(k, so, layout_flat(x, 1), lanes, 1).reshape(lanes, n_out).T
        st = _dec_state(k)(lanes, DEV)
    outs, a = [], 0
    for c in chunks:
        y = torch.empty(lanes * c, dtype=torch.float32, device=DEV)
        Lanes(casc).block(st, to_dev(layout_flat(x[a:a + c], 0)), y, 0)
        if c % TO == 0:
            assert ctx.last_kernel == "hbf tiled frame-major (tensor-map input)", ctx.last_kernel
        outs.append(to_np(y).reshape(c, lanes))
        a += c
    assert_bits_equal(np.concatenate(outs, axis=0), want, f"lanes={lanes} taps98={taps98}")
    assert_bits_equal(st.numpy(), so, "state")


@pytest.mark.parametrize("k", [2, 3, 4, 5])
@pytest.mark.parametrize("lanes", [1, 2, 9, 64, 131])
def test_int16_frame_major_tensor_map_output(oracle, lanes, k):
    """frame-major x16 (`[[f32; 16]; lanes]` output frames): the staged output tile leaves as four tensor-map
    stores per CTA (one lane pair x 32 frames each, 128-byte swizzled staging lines, hbf_int_fast_body.cuh FmOut).
    Odd lane counts (half / whole boxes outside the tensor are clipped), one lane, several CTAs, whole tiles +
    ragged tails, state carried over three calls == the oracle bit for bit."""
    rng = np.random.default_rng(777 + lanes + k)
    R, TI = 1 << k, 512 >> k
    chunks = [3 * TI, 2 * TI + 5, 4 * TI]
    n_in = sum(chunks)
    x = rng.uniform(-1, 1, (n_in, lanes)).astype(np.float32)
    so = np.zeros((oracle.hbf_int_state_words(k), lanes), np.float32)
    want = oracle.hbf_int_cascade_lanes(k, so, layout_flat(x, 1), lanes, 1).reshape(lanes, n_in, R).transpose(1, 0, 2)
    ctx = ib.default_context(0)
    st = _int_state(k)(lanes, DEV)
    outs, a = [], 0
    for c in chunks:
        y = torch.full((lanes * c * R,), float("nan"), dtype=torch.float32, device=DEV)
        Lanes(HbfIntCascade(k)).block(st, to_dev(layout_flat(x[a:a + c], 0)), y, 0)
        if c % TI == 0:
            assert ctx.last_kernel == "hbf tiled frame-major (tensor-map output)", ctx.last_kernel
        outs.append(to_np(y).reshape(c, lanes, R))
        a += c
    assert_bits_equal(np.concatenate(outs, axis=0), want, f"lanes={lanes}")
    assert_bits_equal(st.numpy(), so, "state")
