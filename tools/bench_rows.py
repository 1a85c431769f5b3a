#!/usr/bin/env python3
"""Roofline table over every row of SURVEY.md section 8(a): one timed launch series per
entry point, algorithmic bytes / CUDA-event time vs the measured HBM copy peak.

    python tools/bench_rows.py [--quick]      # writes gpurun_out/rows.json and prints a markdown table
"""
import argparse
import json
import re
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import idsp_b200 as ib  # noqa: E402
from bench import peak_hbm  # noqa: E402
from idsp_b200 import (Accu, Biquad, BiquadClamp, Cascade, DirectForm, DirectForm1, DirectForm1Dither,  # noqa: E402
                       DirectForm1Wide, DirectForm2Transposed, Filter, HbfDecCascade, HbfIntCascade, Lanes, Lockin,
                       LockinState, Lowpass, LowpassState, Q)
from idsp_b200.hbf import _dec_state, _int_state  # noqa: E402

DEV = "cuda:0"
TDT = {"i8": torch.int8, "i16": torch.int16, "i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}
BITS = {"i8": 8, "i16": 16, "i32": 32, "i64": 64}


def rnd(kind, n):
    if kind in BITS:
        b = BITS[kind] - 3
        return torch.randint(-(1 << b), 1 << b, (n,), dtype=torch.int64, device=DEV).to(TDT[kind])
    return torch.empty(n, dtype=TDT[kind], device=DEV).uniform_(-1, 1)


def timeit(fn, reps):
    fn()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None, help="regex: time only the rows whose name matches (e.g. under ncu)")
    ap.add_argument("--reps", type=int, default=None)
    ap.add_argument("--hbf-in", type=int, default=8192, help="high-rate samples per lane and call of the HBF rows")
    ap.add_argument("--policy", type=int, default=0, help="kernel policy (1 = generic kernels only)")
    ap.add_argument("--out", default="gpurun_out/rows.json")
    args = ap.parse_args()
    peak, src = peak_hbm()
    lanes = 65536
    frames = 1024 if args.quick else 4096
    reps = args.reps or (3 if args.quick else 10)
    only = re.compile(args.only) if args.only else None
    rows = []

    def add(name, ref, samples, bytes_per_sample, fn):
        if only and not only.search(name):
            return
        ms = timeit(fn, reps)
        gbs = samples * bytes_per_sample / ms / 1e6
        rows.append({"row": name, "reference": ref, "GSa/s": samples / ms / 1e6, "GB/s": gbs, "frac_of_peak": gbs / peak,
                     "bytes_per_sample": bytes_per_sample, "ms": ms, "kernel": ctx0.last_kernel})
        print(f"{name:46s} {samples / ms / 1e6:9.1f} GSa/s {gbs:8.1f} GB/s {100 * gbs / peak:5.1f}%  [{ctx0.last_kernel}]", flush=True)

    lp = Filter().critical_frequency(0.01).lowpass()
    ctx0 = ib.default_context(0)
    ctx0.set_kernel_policy(args.policy)
    for layout, lname in ((0, "frame-major"), (1, "lane-major")):
        for kind in ("i8", "i16", "i32", "i64", "f32", "f64"):
            fmt = Q(kind, BITS[kind] - 2) if kind in BITS else kind
            bq = Biquad.from_ba6(lp, fmt)
            n = lanes * frames
            x, y = rnd(kind, n), torch.empty(n, dtype=TDT[kind], device=DEV)
            st = DirectForm1.default(kind, lanes, DEV)
            sz = x.element_size()
            add(f"a1/a4 Biquad DF1 {kind} {lname}", "biquad.rs:366-383", n, 2 * sz, lambda: Lanes(bq).block(st, x, y, layout))
            if kind in ("i32", "f32"):
                bc = BiquadClamp(bq, 1, -(1 << 20) if kind == "i32" else -0.5, (1 << 20) if kind == "i32" else 0.5)
                add(f"a2 BiquadClamp DF1 {kind} {lname}", "biquad.rs:394-404", n, 2 * sz, lambda: Lanes(bc).block(st, x, y, layout))
            del x, y
        n = lanes * frames
        x, y = rnd("f32", n), torch.empty(n, dtype=torch.float32, device=DEV)
        bq = Biquad.from_ba6(lp, "f32")
        s2 = DirectForm2Transposed.default("f32", lanes, DEV)
        add(f"a3 Biquad DF2T f32 {lname}", "biquad.rs:418-428", n, 8, lambda: Lanes(bq).block(s2, x, y, layout))
        del x, y
        x, y = rnd("i32", n), torch.empty(n, dtype=torch.int32, device=DEV)
        b29 = Biquad.from_ba6(lp, Q("i32", 29))
        sw, sd = DirectForm1Wide.default(lanes, DEV), DirectForm1Dither.default(lanes, DEV)
        add(f"a5 DirectForm1Wide i32 {lname}", "biquad.rs:445-472", n, 8, lambda: Lanes(b29).block(sw, x, y, layout))
        add(f"a6 DirectForm1Dither i32 {lname}", "biquad.rs:484-530", n, 8, lambda: Lanes(b29).block(sd, x, y, layout))
        casc = Cascade([Biquad.from_ba6(Filter().critical_frequency(0.01 * (i + 1)).lowpass(), Q("i32", 29)) for i in range(4)])
        sc = DirectForm.default(4, "i32", lanes, DEV)
        add(f"a7 Cascade<4> i32 {lname}", "biquad.rs:339-364", n, 8, lambda: Lanes(casc).block(sc, x, y, layout))
        sl1, sl2 = LowpassState.default(1, lanes, DEV), LowpassState.default(2, lanes, DEV)
        add(f"a19 Lowpass<1> i32 {lname}", "lowpass.rs:47-78", n, 8, lambda: Lanes(Lowpass([67465188])).block(sl1, x, y, layout))
        add(f"a19 Lowpass<2> i32 {lname}", "lowpass.rs:47-78", n, 8, lambda: Lanes(Lowpass([1048576, -94906265])).block(sl2, x, y, layout))
        iq = torch.empty(2 * n, dtype=torch.int32, device=DEV)
        acc = Accu(torch.zeros(lanes, dtype=torch.int32, device=DEV), rnd("i32", lanes))
        sk = LockinState.default(2, lanes, DEV)
        add(f"a18-a20 Accu+Lockin<Lowpass<2>> i32 {lname}", "lockin.rs:17-39", n, 12,
            lambda: Lockin(Lowpass([1048576, -94906265])).block(sk, acc, x, iq, layout))
        # caller-supplied phase: (x, phase) pairs in, Complex<i32> out (lockin.rs:30-39)
        xp = rnd("i32", 2 * n)
        add(f"a20 Lockin<Lowpass<2>> on (x, phase) i32 {lname}", "lockin.rs:30-39", n, 16,
            lambda: Lockin(Lowpass([1048576, -94906265])).block_phase(sk, xp, iq, layout))
        del x, y, iq, xp
        # single /2 and x2 stages with run-time taps (a11 / a12) and the 98 dB tap set as a /16 cascade
        from idsp_b200 import EvenSymmetric, HbfDec, HbfInt, hbf_taps, hbf_taps_98
        for ti in (0, 3):
            tp = hbf_taps()[ti]
            sl, sn = 65536, 2048
            xs, ys = rnd("f32", sl * sn * 2), torch.empty(sl * sn, dtype=torch.float32, device=DEV)
            sd_, si_ = HbfDec.default(len(tp), sl, DEV), HbfInt.default(len(tp), sl, DEV)
            add(f"a11 HbfDec /2 single stage M={len(tp)} f32 {lname}", "hbf.rs:155-192", sl * sn * 2, 6,
                lambda: EvenSymmetric(tp).block(sd_, xs, ys, layout))
            add(f"a12 HbfInt x2 single stage M={len(tp)} f32 {lname}", "hbf.rs:207-236", sl * sn * 2, 6,
                lambda: EvenSymmetric(tp).block(si_, ys, xs, layout))
            del xs, ys
        c98 = HbfDecCascade(4, hbf_taps_98()[:4])
        sl, sn = 65536, 256
        xs, ys = rnd("f32", sl * sn * 16), torch.empty(sl * sn, dtype=torch.float32, device=DEV)
        s98 = c98.state(sl, DEV)
        add(f"a13 HbfDec /16 cascade, HBF_TAPS_98 f32 {lname}", "hbf.rs:258-292, 385-421", sl * sn * 16, 4.25,
            lambda: Lanes(c98).block(s98, xs, ys, layout))
        other = HbfDecCascade(4, [t * np.float32(0.5) for t in hbf_taps_98()[:4]])
        so_ = other.state(sl, DEV)
        add(f"a13 HbfDec /16 cascade, run-time taps (stage by stage) f32 {lname}", "hbf.rs:155-192, 385-421", sl * sn * 16, 4.25,
            lambda: Lanes(other).block(so_, xs, ys, layout))
        del xs, ys
        # hbf cascades: 262144 lanes (config 3 shape, fewer frames)
        hl = 65536 if args.quick else 262144
        for k in (1, 2, 3, 4, 5):
            R = 1 << k
            n_out = args.hbf_in >> k  # same high-rate stream length per lane for every rate (state I/O amortised alike)
            x = rnd("f32", hl * n_out * R)
            y = torch.empty(hl * n_out, dtype=torch.float32, device=DEV)
            sdec = _dec_state(k)(hl, DEV)
            add(f"a13 HbfDec /{R} cascade f32 {lname}", "hbf.rs:385-421", hl * n_out * R, 4 + 4 / R,
                lambda: Lanes(HbfDecCascade(k)).block(sdec, x, y, layout))
            sint = _int_state(k)(hl, DEV)
            add(f"a14 HbfInt x{R} cascade f32 {lname}", "hbf.rs:476-512", hl * n_out * R, 4 + 4 / R,
                lambda: Lanes(HbfIntCascade(k)).block(sint, y, x, layout))
            del x, y
        if layout == 1:  # config 5 chain at 65 536 lanes x 16 384 samples
            from idsp_b200 import _lib as _l
            cl, cn = 65536, 1024
            W = int(_l.lib().idsp_chain_state_words(4))
            xc5 = rnd("f32", cl * cn * 16)
            yc5 = torch.empty_like(xc5)
            stc = torch.zeros((W, cl), dtype=torch.float32, device=DEV)
            bac = np.asarray(Biquad.from_ba6(Filter().critical_frequency(0.05).lowpass(), "f32").ba)
            add(f"cfg5 chain HbfDec16->HbfInt16->Biquad f32 {lname}", "hbf.rs:385-512, biquad.rs:366-383", cl * cn * 16, 8,
                lambda: ctx0.chain(4, bac, stc, xc5, yc5, lanes=cl, layout=1))
            del xc5, yc5
        from idsp_b200 import FmDiscriminator, FmDiscState
        xc = rnd("i32", 2 * lanes * frames)
        yd = torch.empty(lanes * frames, dtype=torch.int32, device=DEV)
        sf = FmDiscState.default(lanes, DEV)
        fd = FmDiscriminator(0x19341234, Biquad.from_ba6(lp, Q("i32", 30)))
        add(f"f4 FM discriminator graph i32 {lname}", "examples/fm_disc.rs:26-48", lanes * frames, 12,
            lambda: Lanes(fd).block(sf, xc, yd, layout))
        del xc, yd
        from idsp_b200 import PLL, PLLState
        xp, yp = rnd("i32", lanes * frames), torch.empty(lanes * frames, dtype=torch.int32, device=DEV)
        sp = PLLState.default(lanes, DEV)
        add(f"f4 PLL i32 {lname}", "pll.rs:88-108", lanes * frames, 8,
            lambda: Lanes(PLL.from_bandwidth(1e-2, 4.0)).block(sp, xp, yp, layout))
        del xp, yp
        # CIC /16 and x16, cubic (src/cic.rs PERF_N = 3, PERF_R = 16, PERF_D = 1)
        from idsp_b200 import Cic, CicState
        for kind in ("i32", "i64"):
            cl, cf, R = 65536, (256 if args.quick else 1024), 16
            xh = rnd(kind, cl * cf * R)
            yl = torch.empty(cl * cf, dtype=TDT[kind], device=DEV)
            sz = xh.element_size()
            cd, ci = CicState.default(3, 1, kind, cl, DEV), CicState.default(3, 1, kind, cl, DEV)
            add(f"f3 Cic<3> /16 decimator {kind} {lname}", "cic.rs:176-200", cl * cf * R, sz + sz / R,
                lambda: Lanes(Cic(3, 1, 15).decimate()).block(cd, xh, yl, layout))
            add(f"f3 Cic<3> x16 interpolator {kind} {lname}", "cic.rs:149-172", cl * cf * R, sz + sz / R,
                lambda: Lanes(Cic(3, 1, 15).interpolate()).block(ci, yl, xh, layout))
            del xh, yl
    n = 1 << (24 if args.quick else 28)
    ph = rnd("i32", n)
    cs = torch.empty(2 * n, dtype=torch.int32, device=DEV)
    ctx = ib.default_context(0)
    add("a16 cossin i32", "cossin.rs:14-67", n, 12, lambda: ctx.cossin(ph, cs))
    pp = torch.empty(n, dtype=torch.int32, device=DEV)
    add("a17 atan2 i32", "atan2.rs:66-82", n, 12, lambda: ctx.atan2(cs, pp))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump({"peak_GBs": peak, "peak_source": src, "policy": args.policy, "rows": rows}, open(args.out, "w"), indent=1)
    print("\n| §8 row | reference | GSa/s | algorithmic GB/s | of measured HBM peak |\n|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['row']} | `{r['reference']}` | {r['GSa/s']:.1f} | {r['GB/s']:.0f} | {100 * r['frac_of_peak']:.1f} % |")


if __name__ == "__main__":
    main()
