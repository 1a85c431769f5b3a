#!/usr/bin/env python3
"""i8 frame-major biquad: generic kernel (one lane per thread) vs four lanes packed to a 32-bit word on the
tensor-map kernels, over the lane count (IDSP_I8_PACKED_MIN_LANES selects: a huge value = never packed, 1 = always)."""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import idsp_b200 as ib
from idsp_b200 import Lanes, Biquad, Filter, DirectForm1, Q
bq = Biquad.from_ba6(Filter().critical_frequency(0.01).lowpass(), Q("i8", 6))
ctx = ib.default_context(0)
for lg in (16, 18, 19, 20, 22):
    lanes = 1 << lg
    frames = max((1 << 30) // lanes, 64)
    n = lanes * frames
    x = torch.randint(-100, 100, (n,), dtype=torch.int8, device="cuda")
    y = torch.empty_like(x)
    st = DirectForm1.default("i8", lanes, "cuda")
    Lanes(bq).block(st, x, y, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        Lanes(bq).block(st, x, y, 0)
    e1.record(); torch.cuda.synchronize()
    print(f"  2^{lg} lanes x {frames} frames: {5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:8.1f} GSa/s = {10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:7.1f} GB/s  [{ctx.last_kernel}]  checksum {int(y.to(torch.int64).sum())}", flush=True)
    del x, y
''' % ROOT
for name, v in (("never packed", str(1 << 40)), ("always packed", "1")):
    print(name)
    subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, IDSP_I8_PACKED_MIN_LANES=v), check=False)
