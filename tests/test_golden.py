"""Committed fixtures under tests/golden/:
  reference_kats.json -- known answers transcribed from the reference's own tests (file:line cited)
  vectors.npz         -- seeded inputs + oracle outputs (tests/golden/make_vectors.py)
The CPU tests check the oracle against both; the -m gpu tests push the same vectors through the C
ABI and compare bit patterns."""
import json
import os

import numpy as np
import pytest

from idsp_b200.coefficients import Filter
from idsp_b200.iir import Biquad, Q32

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
VEC = np.load(os.path.join(HERE, "golden", "vectors.npz"))


def _bits(a):
    a = np.asarray(a)
    return a.view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize]) if a.dtype.kind == "f" else a


def test_generator_reproduces_committed_vectors(oracle):
    """the oracle still produces exactly the committed vectors (guards oracle regressions)"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_vectors", os.path.join(HERE, "golden", "make_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = mod.make()
    assert sorted(fresh) == sorted(VEC.files)
    for k in VEC.files:
        assert np.array_equal(_bits(fresh[k]), _bits(VEC[k])), k


def test_reference_kats_on_oracle(oracle):
    k = KATS["biquad_q30_lowpass"]
    ba = Biquad.from_ba6(Filter().critical_frequency(k["f0"]).set_gain(k["gain"]).lowpass(), Q32(30)).ba
    assert list(ba) == k["ba"]
    assert list(oracle.biquad_df1("i32", ba, 30, None, np.zeros(4, np.int32), np.array(k["x"], np.int32))) == k["y"]
    k = KATS["biquad_q30_highpass"]
    ba = Biquad.from_ba6(Filter().critical_frequency(k["f0"]).set_gain(k["gain"]).highpass(), Q32(30)).ba
    assert list(oracle.biquad_df1("i32", ba, 30, None, np.zeros(4, np.int32), np.array(k["x"], np.int32))) == k["y"]
    k = KATS["cossin_zero"]
    assert [int(v) for v in oracle.cossin(np.array([k["phase"]], np.int32))[0]] == [k["cos"], k["sin"]]
    lut = oracle.cossin_table()
    assert [int(v) for v in lut[:4]] == KATS["cossin_lut"]["head"] and int(lut[-1]) == KATS["cossin_lut"]["last"]
    for y, x, want in KATS["atan2_axes"]["cases"]:
        assert int(oracle.atan2(np.array([[x, y]], np.int32))[0]) == want
    c = KATS["cic_unit_rate"]
    assert oracle.cic_gain_log2(c["N"], c["M"], c["rate"]) == c["gain_log2"] and oracle.cic_gain(c["N"], c["M"], c["rate"]) == c["gain"]
    assert oracle.hbf_dec_response_length(4) == KATS["hbf_dec16_response_length"]["value"]
    assert oracle.hbf_int_response_length(4) == KATS["hbf_int16_response_length"]["value"]


@pytest.mark.gpu
def test_vectors_on_gpu():
    """every committed vector through the C ABI on the GPU, bit for bit (outputs and final states)"""
    import torch

    import idsp_b200 as ib
    from idsp_b200 import (PLL, Accu, Cic, CicState, DirectForm1, DirectForm2Transposed, FmDiscriminator, FmDiscState,
                           HbfDecCascade, HbfIntCascade, Lanes, Lockin, LockinState, Lowpass, PLLState)
    from idsp_b200.hbf import _dec_state, _int_state

    dev = "cuda:0"
    d = lambda name: torch.from_numpy(np.ascontiguousarray(VEC[name])).to(dev)

    def check(name, got):
        g = got.cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
        assert np.array_equal(_bits(g.reshape(-1)), _bits(VEC[name].reshape(-1))), name

    lanes = 8
    st = DirectForm1.default("i32", lanes, dev)
    y = torch.empty_like(d("biquad_i32_x"))
    Lanes(Biquad(VEC["biquad_i32_ba"], Q32(30))).block(st, d("biquad_i32_x"), y)
    check("biquad_i32_y", y), check("biquad_i32_state", st.numpy())
    st = DirectForm2Transposed.default("f32", lanes, dev)
    y = torch.empty_like(d("df2t_f32_x"))
    Lanes(Biquad(VEC["df2t_f32_ba"], "f32")).block(st, d("df2t_f32_x"), y)
    check("df2t_f32_y", y), check("df2t_f32_state", st.numpy())
    for k in (1, 4):
        st = _dec_state(k)(4, dev)
        y = torch.empty(VEC[f"hbf_dec{k}_y"].size, dtype=torch.float32, device=dev)
        Lanes(HbfDecCascade(k)).block(st, d(f"hbf_dec{k}_x"), y, 1)
        check(f"hbf_dec{k}_y", y), check(f"hbf_dec{k}_state", st.numpy())
        st = _int_state(k)(4, dev)
        y = torch.empty(VEC[f"hbf_int{k}_y"].size, dtype=torch.float32, device=dev)
        Lanes(HbfIntCascade(k)).block(st, d(f"hbf_int{k}_x"), y, 1)
        check(f"hbf_int{k}_y", y), check(f"hbf_int{k}_state", st.numpy())
    ctx = ib.default_context(0)
    check("cossin_cs", ctx.cossin(d("cossin_phase")))
    check("atan2_p", ctx.atan2(d("atan2_xy").reshape(-1)))
    st = LockinState.default(2, lanes, dev)
    acc = Accu(torch.zeros(lanes, dtype=torch.int32, device=dev), d("lockin_step"))
    iq = torch.empty(2 * VEC["lockin_x"].size, dtype=torch.int32, device=dev)
    Lockin(Lowpass([int(v) for v in VEC["lockin_k"]])).block(st, acc, d("lockin_x"), iq, 0)
    check("lockin_iq", iq), check("lockin_state", st.numpy()), check("lockin_accu", acc.state)
    st = CicState.default(3, 1, "i64", lanes, dev)
    y = torch.empty(VEC["cic_dec_y"].size, dtype=torch.int64, device=dev)
    Lanes(Cic(3, 1, 7).decimate()).block(st, d("cic_dec_x"), y)
    check("cic_dec_y", y), check("cic_dec_state", st.numpy())
    st = PLLState.default(lanes, dev)
    y = torch.empty_like(d("pll_x"))
    Lanes(PLL(VEC["pll_ba"])).block(st, d("pll_x"), y)
    check("pll_y", y), check("pll_state", st.numpy())
    st = FmDiscState.default(lanes, dev)
    y = torch.empty(VEC["fm_y"].size, dtype=torch.int32, device=dev)
    Lanes(FmDiscriminator(0x19341234, Biquad(VEC["biquad_i32_ba"], Q32(30)))).block(st, d("fm_x"), y)
    check("fm_y", y), check("fm_state", st.numpy())
