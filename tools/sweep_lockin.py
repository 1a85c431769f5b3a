#!/usr/bin/env python3
"""Tile-shape / residency sweep of the frame-major lock-in kernel (needs an IDSP_TUNE=1 build selected with
IDSP_B200_LIB): times BASELINE configs[3]'s per-GPU workload (131 072 lanes) for every IDSP_OUT8_CFG.

    IDSP_TUNE=1 python -m idsp_b200.build --out=$PWD/idsp_b200/variants/tune.so
    IDSP_B200_LIB=$PWD/idsp_b200/variants/tune.so python tools/sweep_lockin.py
"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CODE = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np, torch
import oracle as O
from idsp_b200 import Accu, Lockin, LockinState, Lowpass
dev = "cuda:0"
K = [1048576, -94906265]
for lanes, frames in ((131072, 4096), (65536, 4096)):
    n = lanes * frames
    x = torch.randint(-(1 << 30), 1 << 30, (n,), dtype=torch.int32, device=dev)
    iq = torch.empty(2 * n, dtype=torch.int32, device=dev)
    step = torch.randint(-(1 << 31), (1 << 31) - 1, (lanes,), dtype=torch.int64, device=dev).to(torch.int32)
    acc = Accu(torch.zeros(lanes, dtype=torch.int32, device=dev), step)
    st = LockinState.default(2, lanes, dev)
    cfg = Lockin(Lowpass(K))
    cfg.block(st, acc, x, iq, 0)
    torch.cuda.synchronize()
    sub = np.arange(0, lanes, lanes // 16)
    xs = x.view(frames, lanes)[:, sub].contiguous().cpu().numpy().reshape(-1)
    want = O.lockin_lanes(K, np.zeros(sub.size, np.int32), step[sub].cpu().numpy(), np.zeros((4, sub.size), np.int64), xs, sub.size, 0)
    got = iq.view(frames, lanes, 2)[:, sub].contiguous().cpu().numpy().reshape(-1)
    ok = bool(np.array_equal(got, want))
    for _ in range(3):
        cfg.block(st, acc, x, iq, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        cfg.block(st, acc, x, iq, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"cfg={os.environ.get('IDSP_OUT8_CFG','-'):>3} lanes={lanes:7d} {n / ms / 1e6:8.1f} GSa/s parity={'ok' if ok else 'FAIL'}", flush=True)
    del x, iq
''' % ROOT

for cfg in ([None, 0] if os.environ.get("IDSP_SWEEP_ONLY") else [None] + list(range(0, 12))):
    env = dict(os.environ)
    if cfg is None:
        env.pop("IDSP_OUT8_CFG", None)
    else:
        env["IDSP_OUT8_CFG"] = str(cfg)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    if r.returncode:
        sys.stdout.write(f"cfg={cfg} FAILED: {r.stderr[-400:]}\n")
    sys.stdout.flush()
