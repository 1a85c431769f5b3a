#!/usr/bin/env python3
"""Generates tests/golden/vectors.npz: seeded inputs and the CPU oracle's outputs for every hot-path
row at small sizes.  The reference is a Rust crate and cannot be built or imported in this image
(no rustc / cargo), so the committed vectors come from the oracle -- which tests/test_oracle_*.py pin
against the reference's own known-answer tests -- and freeze its behaviour: a later change to the
oracle or to a kernel that alters a single bit fails tests/test_golden.py.

    python tests/golden/make_vectors.py          # rewrites vectors.npz (commit the result)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import oracle as O  # noqa: E402
from idsp_b200.coefficients import Filter  # noqa: E402
from idsp_b200.iir import Biquad, Q32  # noqa: E402


def make():
    rng = np.random.default_rng(20261017)
    v = {}
    lanes, frames = 8, 96
    ba = Biquad.from_ba6(Filter().critical_frequency(0.01).lowpass(), Q32(30)).ba
    x = rng.integers(-(1 << 28), 1 << 28, frames * lanes).astype(np.int32)
    st = np.zeros((4, lanes), np.int32)
    v["biquad_i32_ba"], v["biquad_i32_x"] = ba, x
    v["biquad_i32_y"] = O.biquad_lanes("df1", "i32", ba, 30, None, st, x, lanes, 0)
    v["biquad_i32_state"] = st
    baf = Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), "f32").ba
    xf = rng.standard_normal(frames * lanes).astype(np.float32)
    stf = np.zeros((2, lanes), np.float32)
    v["df2t_f32_ba"], v["df2t_f32_x"] = baf, xf
    v["df2t_f32_y"] = O.biquad_lanes("df2t", "f32", baf, 0, None, stf, xf, lanes, 0)
    v["df2t_f32_state"] = stf
    for k in (1, 4):
        hl, n_out = 4, 72
        xh = rng.uniform(-1, 1, hl * n_out * (1 << k)).astype(np.float32)
        sh = np.zeros((O.hbf_dec_state_words(k), hl), np.float32)
        v[f"hbf_dec{k}_x"] = xh
        v[f"hbf_dec{k}_y"] = O.hbf_dec_cascade_lanes(k, sh, xh, hl, 1)
        v[f"hbf_dec{k}_state"] = sh
        xi = rng.uniform(-1, 1, hl * 40).astype(np.float32)
        si = np.zeros((O.hbf_int_state_words(k), hl), np.float32)
        v[f"hbf_int{k}_x"] = xi
        v[f"hbf_int{k}_y"] = O.hbf_int_cascade_lanes(k, si, xi, hl, 1)
        v[f"hbf_int{k}_state"] = si
    ph = rng.integers(-(1 << 31), 1 << 31, 512).astype(np.int32)
    v["cossin_phase"], v["cossin_cs"] = ph, O.cossin(ph)
    xy = rng.integers(-(1 << 31), 1 << 31, (512, 2)).astype(np.int32)
    v["atan2_xy"], v["atan2_p"] = xy, O.atan2(xy)
    k2 = np.array([1048576, -94906265], np.int32)
    xl = rng.integers(-(1 << 30), 1 << 30, frames * lanes).astype(np.int32)
    a0, step = np.zeros(lanes, np.int32), rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
    lp = np.zeros((4, lanes), np.int64)
    v["lockin_k"], v["lockin_x"], v["lockin_step"] = k2, xl, step
    v["lockin_iq"] = O.lockin_lanes(k2, a0, step, lp, xl, lanes, 0)
    v["lockin_accu"], v["lockin_state"] = a0, lp
    xc = rng.integers(-(1 << 40), 1 << 40, 24 * lanes * 8).astype(np.int64)
    sc = np.zeros((O.cic_state_words(3, 1), lanes), np.int64)
    v["cic_dec_x"] = xc
    v["cic_dec_y"] = O.cic_dec_lanes(3, 1, 7, sc, xc, lanes, 0)
    v["cic_dec_state"] = sc
    pba = O.pll_from_bandwidth(1e-2, 4.0)
    xp = ((np.arange(1, frames + 1, dtype=np.uint64)[:, None] * rng.integers(1, 1 << 32, lanes, dtype=np.uint64)[None, :])
          & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32).reshape(-1)
    sp = np.zeros((9, lanes), np.int32)
    v["pll_ba"], v["pll_x"] = pba, xp
    v["pll_y"] = O.pll_lanes(pba, sp, xp, lanes, 0)
    v["pll_state"] = sp
    xf2 = O.cossin(rng.integers(-(1 << 31), 1 << 31, frames * lanes).astype(np.int32)).reshape(-1)
    sf = np.zeros((7, lanes), np.int32)
    v["fm_x"] = xf2
    v["fm_y"] = O.fm_disc_lanes(0x19341234, ba, 30, sf, xf2, lanes, 0)
    v["fm_state"] = sf
    return v


if __name__ == "__main__":
    O.build()
    np.savez_compressed(os.path.join(HERE, "vectors.npz"), **make())
    print("wrote", os.path.join(HERE, "vectors.npz"), os.path.getsize(os.path.join(HERE, "vectors.npz")), "bytes")
