#!/usr/bin/env python3
"""Tile-shape sweep of the i32 DF1 frame-major TMA kernel (needs an IDSP_TUNE=1 build).

    IDSP_TUNE=1 python -m idsp_b200.build --force && python tools/sweep_biquad.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import BIQUAD_LANES, biquad_coeffs  # noqa: E402
from idsp_b200 import DirectForm1, Lanes  # noqa: E402
from idsp_b200.engine import default_context  # noqa: E402

CFGS = {0: "TF16 S4 O2 WPC1", 1: "TF16 S6 O2 WPC1", 2: "TF32 S3 O2 WPC1", 3: "TF32 S4 O2 WPC1", 4: "TF8 S8 O3 WPC1",
        5: "TF16 S4 O2 WPC2", 6: "TF16 S4 O2 WPC4", 7: "TF16 S4 O2 WIDE2", 8: "TF16 S4 O2 WIDE4", 9: "TF16 S4 O2 WIDE8",
        10: "TF8 S6 O2 WIDE4", 11: "TF32 S3 O2 WIDE2", 12: "TF16 S3 O1 WPC1", 13: "TF16 S6 O3 WIDE4", 14: "TF8 S8 O4 WIDE8",
        15: "TF16 S3 O2 WIDE8", 16: "TF16 S5 O2 WIDE8", 17: "TF32 S3 O2 WIDE8", 18: "TF8 S6 O2 WIDE8", 19: "TF16 S4 O3 WIDE8",
        20: "TF8 S4 O2 WIDE8", 21: "TF4 S8 O4 WIDE8", 22: "TF16 S2 O2 WIDE8"}
if os.environ.get("SWEEP_ONLY"):
    CFGS = {int(k): CFGS[int(k)] for k in os.environ["SWEEP_ONLY"].split(",")}


LM_CFGS = {0: "LM 64B-swizzle 16fr S4 O2", 1: "LM NBOX1 S3 O2", 2: "LM NBOX2 S3 O2", 3: "LM NBOX4 S2 O1", 4: "LM NBOX4 S2 O2",
           5: "LM NBOX2 S2 O2", 6: "LM NBOX1 S4 O2", 7: "LM NBOX8 S2 O1", 8: "LM NBOX2 S4 O2", 9: "LM NBOX2 S3 O1", 10: "LM NBOX3 S3 O2",
           11: "LM NBOX3 S2 O1", 12: "LM NBOX3 S3 O1"}


def main():
    layout = 1 if os.environ.get("SWEEP_LM") else 0
    global CFGS
    if layout == 1:
        CFGS = LM_CFGS
    dev = "cuda:0"
    ctx = default_context(0)
    lanes, frames = BIQUAD_LANES, 16384
    n = lanes * frames
    x = [torch.randint(-(1 << 28), 1 << 28, (n,), dtype=torch.int32, device=dev) for _ in range(2)]
    y = torch.empty(n, dtype=torch.int32, device=dev)
    cfg = Lanes(biquad_coeffs())
    st = DirectForm1.default("i32", lanes, dev)
    res = {}
    ref = None
    for c, name in list(CFGS.items()) + [(-1, "generic LDG (policy 1)")]:
        if c >= 0:
            os.environ["IDSP_TMA_LMCFG" if layout else "IDSP_TMA_CFG"] = str(c)
            ctx.set_kernel_policy(0)
        else:
            ctx.set_kernel_policy(1)
        st.words.zero_()
        cfg.block(st, x[0], y, layout)
        torch.cuda.synchronize()
        chk = int(y.sum(dtype=torch.int64).item())
        if ref is None:
            ref = chk
        for i in range(3):
            cfg.block(st, x[i % 2], y, layout)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 20
        for i in range(K):
            cfg.block(st, x[i % 2], y, layout)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        res[name] = {"ms": ms, "GSa/s": n / ms / 1e6, "GB/s": 8 * n / ms / 1e6, "checksum_ok": chk == ref}
        print(f"{name:28s} {ms:8.3f} ms  {n / ms / 1e6:8.1f} GSa/s  {8 * n / ms / 1e6:8.1f} GB/s  checksum_ok={chk == ref}", flush=True)
    json.dump(res, open(os.path.join("gpurun_out", "sweep_biquad_lm.json" if layout else "sweep_biquad.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
