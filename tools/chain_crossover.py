#!/usr/bin/env python3
"""Chain (configs[4]) above 2^17 lanes: single-pass thread-per-lane kernel (default) vs the two-pass tiled path
(kernel policy 2), ~2^30 samples per point."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import idsp_b200 as ib
from idsp_b200 import _lib
from idsp_b200.engine import default_context

ctx = default_context(0)
k = 4
ba = np.asarray(ib.Biquad.from_ba6(ib.Filter().critical_frequency(0.05).lowpass(), "f32").ba)
W = int(_lib.lib().idsp_chain_state_words(k))
for lg in (17, 18, 20, 22):
    lanes = 1 << lg
    n_low = max((1 << 30) // lanes // 16, 32)
    n = lanes * n_low * 16
    x = torch.empty(n, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    y = torch.empty_like(x)
    st = torch.zeros((W, lanes), dtype=torch.float32, device="cuda")
    ref = None
    for policy in (0, 2):
        ctx.set_kernel_policy(policy)
        st.zero_()
        ctx.chain(k, ba, st, x, y, lanes=lanes, layout=1)
        torch.cuda.synchronize()
        got = y[: 1 << 20].clone()
        if ref is None:
            ref = got
        same = bool(torch.equal(ref.view(torch.int32), got.view(torch.int32)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ctx.chain(k, ba, st, x, y, lanes=lanes, layout=1)
        e1.record()
        torch.cuda.synchronize()
        print(f"2^{lg} lanes x {n_low * 16} samples  policy {policy}: {3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:7.1f} GSa/s  [{ctx.last_kernel}] same bits: {same}", flush=True)
    ctx.set_kernel_policy(0)
    del x, y, st
    torch.cuda.empty_cache()
