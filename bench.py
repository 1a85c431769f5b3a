#!/usr/bin/env python3
"""bench.py -- headline benchmark of the lane engine (BASELINE.json metric: GSa/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload biquad|hbf|lockin|chain|plumbing] [--full] [--impl reference]

Workloads (BASELINE.json `configs`):
  biquad (default, configs[1]): 65 536-lane i32 iir::Biquad DF1, Q30 Butterworth lowpass
      f0 = 0.01, shared coefficients (dsp_process::Lanes), frame-major.  The full job is
      1e7 frames per lane (2.6 TB in + 2.6 TB out), far beyond HBM, so it is streamed: one
      STEP = one resident block of `--frames` frames (default 16 384 = 4 GiB in + 4 GiB out),
      filter state carried from step to step, inputs cycled through a ring of distinct
      blocks (each >> the 126 MB L2).  `--full` runs all ceil(1e7/frames) steps.
  hbf (configs[2]): HbfDec /16 cascade, f32, 262 144 lanes x 65 536 inputs per lane,
      lane-major, processed as `--hbf-slices` lane slices per step.

`value`  : GSa/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`    : same metric through the C ABI `*_host` entry point with pinned HOST buffers
           (H2D + kernel + D2H inside the timed region).
`roofline`: algorithmic bytes per launch / average launch duration vs MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (a C port of the reference; the Rust crate cannot be built in
           this image) timed on the host cores on a bounded sample of the same workload.
`extra`  : the default run (any N) also measures configs[2] (hbf, lane-major with all legs and frame-major
           `[[f32;16]; lanes]` frames resident), configs[3] (lock-in, resident per GPU AND as the
           sharded data plane: root-resident lanes -> NCCL scatter -> lock-in -> NCCL gather / kernels storing
           into the root's buffer over NVLink) and configs[4] (chain sweep) in the same process and attaches
           their lines (a failure there is recorded, never raised).
`--impl reference`: times that CPU path alone and prints the same JSON line with impl=reference.
N > 1: one process per GPU (torchrun), lanes sharded, no data-path collective ("weak" scaling:
every rank runs the full 65 536-lane workload on its own lane block); `extra.lockin_sharded` is the
one leg with real data movement between GPUs (the edges of the sharded job, BASELINE configs[3]).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_FRAMES = 10_000_000
BIQUAD_LANES = 65_536
HBF_LANES = 262_144
HBF_INPUTS = 65_536
F_BITS = 30


def biquad_coeffs():
    from idsp_b200.coefficients import Filter
    from idsp_b200.iir import Biquad, Q32

    return Biquad.from_ba6(Filter().critical_frequency(0.01).lowpass(), Q32(F_BITS))


def host_threads():
    """host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so do not ask OpenMP)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(key):
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


def strided_lanes(lanes, n=64):
    """lane subset of the parity gates: spread over the WHOLE lane range (every CTA / TMA box of the launch is
    sampled somewhere) plus the first and the last 8 lanes (the last, possibly ragged, box)"""
    a = np.linspace(0, lanes - 1, max(n - 16, 2)).astype(np.int64)
    return np.unique(np.concatenate([np.arange(min(8, lanes)), a, np.arange(max(lanes - 8, 0), lanes)]))


def pcie_peak(dev, mb=512, reps=8, world=1, bidir=True):
    """pinned-memory copy rates with both directions busy at once (what the e2e leg is bound by):
    returns (h2d GB/s, d2h GB/s) measured with CUDA events on two streams.  At N > 1 every rank runs it
    at the same time (barrier before every repetition): the GPUs share the host's memory system and
    PCIe root, so the yardstick must be measured under the same contention as the e2e leg."""
    import torch

    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = [0.0, 0.0]
    for _ in range(reps + 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        barrier(world)
        with torch.cuda.stream(s1):
            ev[0].record()
            d_in.copy_(h_in, non_blocking=True)
            ev[1].record()
        with torch.cuda.stream(s2):
            ev[2].record()
            if bidir:  # bidir=False: host -> device alone (a decimator returns 1/16 of what it reads)
                h_out.copy_(d_out, non_blocking=True)
            else:
                h_out[: n // 16].copy_(d_out[: n // 16], non_blocking=True)
            ev[3].record()
        torch.cuda.synchronize()
        best[0] = max(best[0], n / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9)
        best[1] = max(best[1], n / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9)
    if world > 1:  # the slowest rank bounds the job
        import torch.distributed as dist

        t = torch.tensor(best, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        best = [float(t[0]), float(t[1])]
    return best[0], best[1]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU runs the timed workload.  A reader
    thread collects the lines, so the caller can keep the GPU busy until enough samples exist
    (nvidia-smi needs ~0.1-0.5 s to start on an 8-GPU box, longer than a short timed region)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        import threading

        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def reader():
            for line in self.proc.stdout:
                self.lines.append(line)

        self.thread = threading.Thread(target=reader, daemon=True)
        self.thread.start()

    def count(self):
        return len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def keep_busy_until_sampled(sampler, step, min_samples=3, max_seconds=2.0):
    """after the timed region: keep running (untimed) steps of the same workload until the clock
    sampler has seen the GPU under this load at least `min_samples` times"""
    import torch

    t0 = time.perf_counter()
    i = 0
    while sampler.proc and sampler.count() < min_samples and time.perf_counter() - t0 < max_seconds:
        for _ in range(8):
            step(i)
            i += 1
        torch.cuda.synchronize()


def dist_setup(n_gpus):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        # NCCL writes its version banner to stdout when the communicator is created; stdout must
        # carry the one JSON line only, so fd 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        torch.cuda.set_device(local)
    return rank, world, local


def max_over_ranks(ms, world, dev):
    import torch

    if world == 1:
        return ms
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


# ----------------------------------------------------------------------------- CPU arm
def cpu_biquad(frames, threads, seconds=0.0):
    """oracle DF1 i32 frame-major on `threads` host threads over a resident block of
    `frames` frames, repeated (state carried, like the GPU steps) for >= `seconds`;
    returns (GSa/s, elapsed s, frames processed)"""
    import oracle as O

    bq = biquad_coeffs()
    rng = np.random.default_rng(2)
    x = rng.integers(-(1 << 28), 1 << 28, frames * BIQUAD_LANES, dtype=np.int64).astype(np.int32)
    st = np.zeros((4, BIQUAD_LANES), np.int32)
    O.biquad_lanes("df1", "i32", bq.ba, F_BITS, None, st, x[: 8 * BIQUAD_LANES], BIQUAD_LANES, 0, nthreads=threads)
    t0 = time.perf_counter()
    done = 0
    while True:
        O.biquad_lanes("df1", "i32", bq.ba, F_BITS, None, st, x, BIQUAD_LANES, 0, nthreads=threads)
        done += frames
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return done * BIQUAD_LANES / dt / 1e9, dt, done


def cpu_hbf(lanes, n_out, threads, seconds=0.0):
    import oracle as O

    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, lanes * n_out * 16).astype(np.float32)
    st = np.zeros((O.hbf_dec_state_words(4), lanes), np.float32)
    t0 = time.perf_counter()
    done = 0
    while True:
        O.hbf_dec_cascade_lanes(4, st, x, lanes, 1, nthreads=threads)
        done += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return done * lanes * n_out * 16 / dt / 1e9, dt, done


def cpu_calibrated(workload, threads, target_s):
    """bounded sample: a resident block processed repeatedly for ~target_s seconds"""
    import oracle as O

    O.build_native()
    if workload == "biquad":
        g, dt, frames = cpu_biquad(2048, threads, target_s)
        return g, dt, f"{BIQUAD_LANES} lanes x {frames} frames (i32 DF1, frame-major, 2048-frame resident block repeated)"
    g, dt, reps = cpu_hbf(16384, 1024, threads, target_s)
    return g, dt, f"16384 lanes x {1024 * 16} inputs x {reps} passes (f32 HbfDec/16, lane-major)"


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; Rust crate unbuildable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O

    O.build_native()  # -march=native on the host it is timed on (falls back to the portable build)
    threads = host_threads()
    wl = args.workload
    # The SAME config as the GPU arm (same lanes, same frames_per_step, same dict).  A step's frames are
    # processed as a resident block of `blk` frames passed frames/blk times with the filter state carried --
    # the per-sample work and the state handling are those of the full step, the input block (>> the CPU's
    # caches: 512 MiB) is re-read instead of being generated 4 GiB at a time.
    if wl == "biquad":
        import oracle as O2

        bq = biquad_coeffs()
        frames, blk = args.frames, min(args.frames, 2048)
        rng = np.random.default_rng(2)
        x = rng.integers(-(1 << 28), 1 << 28, blk * BIQUAD_LANES, dtype=np.int64).astype(np.int32)
        st = np.zeros((4, BIQUAD_LANES), np.int32)
        passes = max(1, frames // blk)

        def step():
            t0 = time.perf_counter()
            for _ in range(passes):
                O2.biquad_lanes("df1", "i32", bq.ba, F_BITS, None, st, x, BIQUAD_LANES, 0, nthreads=threads)
            return time.perf_counter() - t0

        samples = passes * blk * BIQUAD_LANES
        sample = f"{BIQUAD_LANES} lanes x {passes * blk} frames per step ({passes} passes over a resident {blk}-frame block, state carried)"
        cfg = biquad_config(args, frames)
    else:
        # one step = one lane slice of the job, like the GPU arm (HBF_LANES / slices lanes x 65536 inputs),
        # processed as passes over a resident block of 2048 lanes x 65536 inputs (512 MiB)
        lanes_s, blk_l = HBF_LANES // args.hbf_slices, 2048
        passes = max(1, lanes_s // blk_l)
        step = lambda: sum(cpu_hbf(blk_l, HBF_INPUTS // 16, threads)[1] for _ in range(passes))  # noqa: E731
        samples = passes * blk_l * HBF_INPUTS
        sample = f"{passes * blk_l} lanes x {HBF_INPUTS} inputs per step ({passes} passes over a resident {blk_l}-lane block)"
        cfg = hbf_config(args)
    for _ in range(args.warmup):
        step()
    t = 0.0
    for _ in range(args.steps):
        t += step()
    val = samples * args.steps / t / 1e9
    line = {
        "impl": "reference", "metric": metric_name(wl), "value": val, "unit": "GSa/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32" if wl == "biquad" else "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "GSa/s", "cores": threads, "kind": "port", "sample": sample,
                         "build": O.build_flags() + " -ffp-contract=off -fopenmp"},
        "e2e": {"value": val, "unit": "GSa/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = C port of the Rust crate's scalar loops (oracle/), all host threads; "
                "rustc/cargo are not in this image so the crate itself cannot be built",
    }
    print(json.dumps(line), flush=True)


def metric_name(wl):
    return ("GSa/s, 65536-lane i32 iir::Biquad DF1 (dsp_process::Lanes), per-step resident block"
            if wl == "biquad" else "GSa/s, f32 hbf::HbfDec /16 cascade, 262144 lanes x 65536 inputs")


def biquad_config(args, frames):
    return {"workload": "configs[1]: 65536-lane i32 iir::Biquad DF1, shared Q30 lowpass f0=0.01, frame-major",
            "lanes_per_gpu": BIQUAD_LANES, "frames_per_step": frames, "total_frames_of_full_job": TOTAL_FRAMES,
            "streaming": "state carried across steps; ring of distinct input blocks, each >> L2",
            "parallelism": f"lanes sharded over {args.gpus} GPU(s), no collective"}


def hbf_config(args):
    return {"workload": "configs[2]: hbf::HbfDec /16 (TAPS.3->2->1->0), f32, 262144 lanes x 65536 inputs, lane-major",
            "lanes_per_gpu": HBF_LANES, "inputs_per_lane": HBF_INPUTS, "slices_per_job": args.hbf_slices,
            "l2": "inputs (64 GiB) >> L2", "parallelism": f"lanes sharded over {args.gpus} GPU(s), no collective"}


# ----------------------------------------------------------------------------- GPU arm
def time_steps(step, steps, warmup, world, local, ctx, dev):
    """W warm-up steps, then exactly K timed steps between barrier+sync, CUDA events on the
    launching stream, max over ranks; returns (ms, launches, clocks)."""
    import torch

    for i in range(warmup):
        step(i)
    sampler = ClockSampler(local)
    barrier(world)
    l0 = ctx.launches
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    launches = ctx.launches - l0
    keep_busy_until_sampled(sampler, step)
    clocks = sampler.stop()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world, dev), launches, clocks


def sustained_loop(step, ms_per_step, world, dev, seconds=2.0):
    """the same step repeated for >= `seconds` (the timed K steps are a short burst at boost clocks; a long
    job runs under the power cap): returns (ms per step, steps)"""
    import torch

    n = int(max(8, min(20000, seconds * 1e3 / max(ms_per_step, 1e-3))))
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    return max_over_ranks(e0.elapsed_time(e1), world, dev) / n, n


LOCKIN_LANES_TOTAL = 1_048_576
LOCKIN_FRAMES = 16_384
LOCKIN_K = [1048576, -94906265]  # Lowpass<2>: [k^2/2^32, -k/q], k = 2^26, q = 1/sqrt(2)


def run_lockin(args, rank, world, local):
    """configs[3]: DDC lock-in (Accu + cossin NCO -> mix -> Lockin<Lowpass<2>>), i32, 1 048 576 lanes
    sharded over 8 GPUs = 131 072 lanes per GPU (kept per GPU for any N: weak scaling)."""
    import torch

    import oracle as O
    from idsp_b200 import Accu, Lockin, LockinState, Lowpass
    from idsp_b200.engine import default_context

    dev = f"cuda:{local}"
    ctx = default_context(local)
    lanes, frames = LOCKIN_LANES_TOTAL // 8, LOCKIN_FRAMES
    n = lanes * frames
    gen = torch.Generator(device=dev)
    gen.manual_seed(4 + rank)
    xin = [torch.randint(-(1 << 30), 1 << 30, (n,), dtype=torch.int32, device=dev, generator=gen) for _ in range(2)]
    iq = [torch.empty(2 * n, dtype=torch.int32, device=dev) for _ in range(2)]
    step_t = torch.randint(-(1 << 31), (1 << 31) - 1, (lanes,), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
    acc = Accu(torch.zeros(lanes, dtype=torch.int32, device=dev), step_t)
    st = LockinState.default(2, lanes, dev)
    cfg = Lockin(Lowpass(LOCKIN_K))

    def step(i):
        cfg.block(st, acc, xin[i % 2], iq[i % 2], 0)

    step(0)
    torch.cuda.synchronize()
    idx = torch.from_numpy(strided_lanes(lanes, 48)).to(dev)
    sub = int(idx.numel())
    xs = xin[0].view(frames, lanes).index_select(1, idx).contiguous().cpu().numpy().reshape(-1)
    a0 = np.zeros(sub, np.int32)
    so = np.zeros((4, sub), np.int64)
    want = O.lockin_lanes(LOCKIN_K, a0, step_t.index_select(0, idx).cpu().numpy(), so, xs, sub, 0, nthreads=host_threads())
    got = iq[0].view(frames, lanes, 2).index_select(1, idx).contiguous().cpu().numpy().reshape(-1)
    if not np.array_equal(got, want):
        raise SystemExit("bench: GPU lock-in output differs from the oracle -- refusing to report a number")
    ms, launches, clocks = time_steps(step, args.steps, args.warmup, world, local, ctx, dev)
    value = world * n * args.steps / (ms * 1e-3) / 1e9
    if rank != 0:
        return None
    peak, peak_src = peak_hbm()
    per_launch_bytes = 12.0 * n
    achieved = per_launch_bytes * launches / (ms * 1e-3) / 1e9 if launches else 0.0
    return {
        "metric": "GSa/s, i32 DDC lock-in (Accu+cossin -> mix -> Lockin<Lowpass<2>>), 131072 lanes per GPU", "value": value,
        "unit": "GSa/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": {"workload": "configs[3]: DDC lock-in i32, 1048576 lanes over 8 GPUs (131072 per GPU), 16384 frames, frame-major",
                   "lanes_per_gpu": lanes, "frames_per_step": frames, "parallelism": f"lanes sharded over {world} GPU(s), no collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic_from_profiles("lockin_i32_fm_bytes_per_launch") if (lanes, frames) == (131072, 16384) else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch_bytes,
                     "note": "52 integer instructions per sample; ALU pipe 60 %, IMAD (fmaheavy) pipe 66 % busy (profiles/r2_lockin_fm_final_ncu.md): bound by the two integer pipes, not by HBM"},
        "gpu_launches": int(launches), "clocks": clocks,
        "parity_check": f"first step == oracle on {sub} lanes strided over the whole lane range (+ first / last 8) x all frames",
    }


def run_lockin_sharded(args, rank, world, local):
    """configs[3] as a DATA PLANE: rank 0 holds the samples of all `131072 x world` lanes (1 048 576 at 8 GPUs),
    contiguous lane blocks go to the ranks (idsp_scatter_lanes: raw NCCL send/recv over NVLink), every rank runs
    the fused lock-in on its own block (no collective inside the computation, compose.rs:472-475), and the
    Complex<i32> results return to rank 0 either (a) through idsp_gather_lanes or (b) straight from the kernels'
    epilogues into rank 0's buffer mapped over NVLink (CUDA IPC peer memory).  Lane-major, so that a lane block
    of the root's buffers is one contiguous range.  Each phase is timed with CUDA events on every rank (max over
    ranks); the gathered result is compared with the CPU oracle on lanes strided over every rank's block."""
    import torch

    import oracle as O
    from idsp_b200 import Accu, Lockin, LockinState, Lowpass
    from idsp_b200.dist import Comm, PeerBuffer
    from idsp_b200.engine import default_context

    dev = f"cuda:{local}"
    ctx = default_context(local)
    lpg = LOCKIN_LANES_TOTAL // 8
    lanes, frames = lpg * world, args.sharded_frames
    comm = Comm(local)
    lo, hi = comm.lane_block(lanes)
    nl = hi - lo
    gen = torch.Generator(device=dev)
    gen.manual_seed(44)
    x_full = step_full = None
    if rank == 0:
        x_full = torch.randint(-(1 << 30), 1 << 30, (lanes * frames,), dtype=torch.int32, device=dev, generator=gen)
        step_full = torch.randint(-(1 << 31), (1 << 31) - 1, (lanes,), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
    # the root's result buffer: a plain device allocation exported to every rank (also the gather target)
    pb = PeerBuffer(lanes * frames * 2, torch.int32, local, owner=0) if world > 1 else None
    if rank == 0:
        iq_full = pb.tensor() if pb is not None else torch.empty(lanes * frames * 2, dtype=torch.int32, device=dev)
    else:
        iq_full = None
    xs = torch.empty(nl * frames, dtype=torch.int32, device=dev)
    iq = torch.empty(nl * frames * 2, dtype=torch.int32, device=dev)
    steps_l = comm.scatter_lanes(step_full, 1, lanes, 1, dtype=torch.int32)
    cfg = Lockin(Lowpass(LOCKIN_K))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def one_pass(fused):
        """scatter -> lock-in -> gather (or lock-in storing into the root's buffer); returns per-phase ms"""
        st = LockinState.default(2, nl, dev)
        acc = Accu(torch.zeros(nl, dtype=torch.int32, device=dev), steps_l)
        e = [ev() for _ in range(4)]
        barrier(world)
        e[0].record()
        comm.scatter_lanes(x_full, frames, lanes, 1, out=xs)
        e[1].record()
        if fused and world > 1:
            cfg.block(st, acc, xs, pb.view(lo * frames * 2, nl * frames * 2), 1)
            e[2].record()
            e[3].record()
        else:
            cfg.block(st, acc, xs, iq, 1)
            e[2].record()
            comm.gather_lanes(iq, iq_full, frames, lanes, 1, width=2)
            e[3].record()
        torch.cuda.synchronize()
        barrier(world)
        t = [e[i].elapsed_time(e[i + 1]) for i in range(3)] + [e[0].elapsed_time(e[3])]
        return [max_over_ranks(v, world, dev) for v in t]

    def check(tag):
        if rank != 0:
            return True
        # lanes strided over the block of every rank (first, last and a few inside)
        sel = []
        for r in range(world):
            a, b = r * lpg, (r + 1) * lpg
            sel += [a, a + 1, a + lpg // 3, a + lpg // 2 + 17, b - 2, b - 1]
        idx = torch.tensor(sel, dtype=torch.int64, device=dev)
        xsub = x_full.view(lanes, frames).index_select(0, idx).contiguous().cpu().numpy().reshape(-1)
        want = O.lockin_lanes(LOCKIN_K, np.zeros(len(sel), np.int32), step_full.index_select(0, idx).cpu().numpy(),
                              np.zeros((4, len(sel)), np.int64), xsub, len(sel), 1, nthreads=host_threads())
        got = iq_full.view(lanes, frames * 2).index_select(0, idx).contiguous().cpu().numpy().reshape(-1)
        return bool(np.array_equal(got, want))

    NSUB = 4  # sub-blocks of a rank's lane block in the pipelined variant
    sub_l = nl // NSUB
    comm_p = Comm(local, own_stream=True) if world > 1 else None  # transfers on their own stream
    xs_sub = [xs[j * sub_l * frames:(j + 1) * sub_l * frames] for j in range(NSUB)]

    def one_pass_pipelined():
        """(c) the block travels in NSUB sub-blocks: sub-block j+1 is on the wire (root egress) while sub-block j is
        filtered and its result tiles are stored into the root's buffer (root ingress) -- both NVLink directions
        busy at once.  The root filters its own lanes in place.  Returns total ms (max over ranks)."""
        sts = [LockinState.default(2, sub_l, dev) for _ in range(NSUB)]
        accs = [Accu(torch.zeros(sub_l, dtype=torch.int32, device=dev), steps_l[j * sub_l:(j + 1) * sub_l].contiguous()) for j in range(NSUB)]
        e0, e1 = ev(), ev()
        barrier(world)
        e0.record()
        comm_p.after_compute()  # the transfers start after e0
        for j in range(NSUB):
            if rank == 0:
                with comm_p.group():
                    for p in range(1, world):
                        a = (p * lpg + j * sub_l) * frames
                        comm_p.send(x_full[a:a + sub_l * frames], p)
                xj = x_full[(lo + j * sub_l) * frames:(lo + (j + 1) * sub_l) * frames]
                out = iq_full[(lo + j * sub_l) * frames * 2:(lo + (j + 1) * sub_l) * frames * 2]
            else:
                comm_p.recv(xs_sub[j], 0)
                comm_p.compute_after()  # the kernel on sub-block j waits for its samples only
                xj, out = xs_sub[j], pb.view((lo + j * sub_l) * frames * 2, sub_l * frames * 2)
            cfg.block(sts[j], accs[j], xj, out, 1)
        e1.record()
        torch.cuda.synchronize()
        comm_p.sync()
        barrier(world)
        return max_over_ranks(e0.elapsed_time(e1), world, dev)

    res = {}
    for fused in ((False, True) if world > 1 else (False,)):
        if rank == 0:
            iq_full.zero_()
        one_pass(fused)  # warm-up (NCCL channels, peer mappings) -- and the pass the oracle checks
        ok = check("fused" if fused else "nccl")
        okt = torch.tensor([1 if ok else 0], device=dev)
        if world > 1:
            import torch.distributed as dist

            dist.broadcast(okt, 0)
        if not int(okt.item()):
            raise SystemExit("bench: sharded lock-in output differs from the oracle -- refusing to report a number")
        reps = max(2, min(args.steps, 5))
        tt = np.array([one_pass(fused) for _ in range(reps)])
        res["fused" if fused else "nccl"] = tt.mean(0)
    if world > 1 and nl % (NSUB * 32) == 0:
        if rank == 0:
            iq_full.zero_()
        one_pass_pipelined()
        okt = torch.tensor([1 if check("pipelined") else 0], device=dev)
        import torch.distributed as dist

        dist.broadcast(okt, 0)
        if not int(okt.item()):
            raise SystemExit("bench: pipelined sharded lock-in output differs from the oracle -- refusing to report a number")
        res["pipelined"] = float(np.mean([one_pass_pipelined() for _ in range(max(2, min(args.steps, 5)))]))
    if comm_p is not None:
        comm_p.close()
    comm.close()
    if pb is not None:
        del iq_full
        pb.close()
    if rank != 0:
        return None
    n = lanes * frames
    t = res["nccl"]
    out_bytes_remote = 8.0 * n * (world - 1) / world
    in_bytes_remote = 4.0 * n * (world - 1) / world
    line = {
        "metric": "GSa/s, i32 DDC lock-in, lanes sharded from / to rank 0 (scatter -> lock-in -> gather over NVLink)",
        "unit": "GSa/s", "n_gpus": world, "higher_is_better": True, "scaling": "weak", "dtype": "i32", "data": "synthetic",
        "config": {"workload": f"configs[3]: DDC lock-in i32, {lanes} lanes sharded over {world} GPU(s) ({lpg} per GPU), "
                               f"{frames} of the 16384 frames resident on the root per pass, lane-major",
                   "lanes_total": lanes, "lanes_per_gpu": lpg, "frames_per_pass": frames,
                   "edges": "idsp_scatter_lanes / idsp_gather_lanes (C ABI, raw NCCL send/recv, in place for lane-major blocks)"},
        "value": n / (t[3] * 1e-3) / 1e9,  # end to end over NVLink: scatter + compute + gather
        "resident_GSa/s": n / (t[1] * 1e-3) / 1e9,  # the lock-in kernels alone (all ranks, max over ranks)
        "ms": {"scatter": t[0], "lockin": t[1], "gather": t[2], "total": t[3]},
        "nvlink_GBs": {"scatter_egress_of_root": (in_bytes_remote / (t[0] * 1e-3) / 1e9) if world > 1 else None,
                       "gather_ingress_of_root": (out_bytes_remote / (t[2] * 1e-3) / 1e9) if world > 1 else None,
                       "nominal_per_direction": 900.0},
        "parity_check": f"gathered Complex<i32> of 6 lanes per rank block ({6 * world} lanes, first / last / inside) x all frames == oracle, both variants",
        "passes_timed": int(max(2, min(args.steps, 5))),
    }
    if "fused" in res:
        f = res["fused"]
        line["fused_peer_store"] = {
            "what": "kernels store their result tiles into rank 0's buffer over NVLink from their own epilogue (no separate gather)",
            "value": n / (f[3] * 1e-3) / 1e9, "ms": {"scatter": f[0], "lockin_with_stores_to_root": f[1], "total": f[3]},
            "ingress_of_root_GBs": out_bytes_remote / (f[1] * 1e-3) / 1e9,
        }
    if "pipelined" in res:
        pm = res["pipelined"]
        line["pipelined_scatter_compute_store"] = {
            "what": f"the lane block travels in {NSUB} sub-blocks (idsp_comm_send / _recv on the communicator's own stream): "
                    "sub-block j+1 is on the wire while sub-block j is filtered and its tiles are stored into rank 0's buffer, "
                    "so root egress and root ingress overlap",
            "value": n / (pm * 1e-3) / 1e9, "ms_total": pm, "sub_blocks": NSUB,
            "root_link_GBs": {"egress": in_bytes_remote / (pm * 1e-3) / 1e9, "ingress": out_bytes_remote / (pm * 1e-3) / 1e9},
        }
    return line


def run_chain(args, rank, world, local):
    """configs[4]: HbfDec(/16) -> HbfInt(x16) -> Biquad DF1 f32 chain, lane sweep 2^10..2^24 at
    ~2^30 samples per point; reports GSa/s per lane count."""
    import torch

    import oracle as O
    from idsp_b200.engine import default_context

    dev = f"cuda:{local}"
    ctx = default_context(local)
    k = 4
    ba = np.asarray(__import__("idsp_b200").Biquad.from_ba6(__import__("idsp_b200").Filter().critical_frequency(0.05).lowpass(), "f32").ba)
    W = int(__import__("idsp_b200")._lib.lib().idsp_chain_state_words(k))  # product state size from the ABI; the oracle below only checks
    total = 1 << 30
    points = []
    for lg in range(10, 25, 2):
        lanes = 1 << lg
        n_low = max(total // lanes // 16, 32)
        n = lanes * n_low * 16
        x = torch.empty(n, dtype=torch.float32, device=dev).uniform_(-1, 1)
        y = torch.empty_like(x)
        st = torch.zeros((W, lanes), dtype=torch.float32, device=dev)

        def step(i):
            ctx.chain(k, ba, st, x, y, lanes=lanes, layout=1)

        if lg in (10, 16, 20):  # one point per kernel path: fused two-pass (8- and 16-lane tiles), single pass
            step(0)
            torch.cuda.synchronize()
            idx = torch.from_numpy(strided_lanes(lanes, 24)).to(dev)
            sub, ns = int(idx.numel()), min(n_low, 512) * 16
            xs = x.view(lanes, n_low * 16).index_select(0, idx)[:, :ns].contiguous().cpu().numpy()
            so = np.zeros((W, sub), np.float32)
            want = O.chain_lanes(k, ba, so, xs.reshape(-1), sub, 1, nthreads=host_threads())
            got = y.view(lanes, n_low * 16).index_select(0, idx)[:, :ns].contiguous().cpu().numpy().reshape(-1)
            if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
                raise SystemExit(f"bench: GPU chain output differs from the oracle at 2^{lg} lanes")
            st.zero_()
        ms, launches, clocks = time_steps(step, 3, 2, world, local, ctx, dev)
        # the biquad recurrence advances one sample per lane per dependent FMUL -> FADD -> FADD chain (3 x 4 cycles on
        # sm_100a, bit-exactness forbids re-association): with few lanes that chain, not bandwidth, bounds the job
        sm_ghz = ((clocks or {}).get("sm_mhz") or 1900.0) / 1e3
        points.append({"lanes_per_gpu": lanes, "samples_per_lane": n_low * 16, "GSa/s": world * n * 3 / (ms * 1e-3) / 1e9,
                       "GB/s": 8.0 * n * 3 / (ms * 1e-3) / 1e9, "kernel": ctx.last_kernel,
                       "recurrence_critical_path_bound_GSa/s": world * lanes * sm_ghz / 12.0})
        del x, y, st
        torch.cuda.empty_cache()
    # end to end through ONE host call (idsp_chain_f32_host): the three operators share one PCIe round trip
    el, en = 65536, 128
    xh = torch.empty(el * en * 16, dtype=torch.float32).uniform_(-1, 1).pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    sth = np.zeros((W, el), np.float32)
    ctx.chain(k, ba, sth, xh.numpy(), yh.numpy(), lanes=el, layout=1)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.chain(k, ba, sth, xh.numpy(), yh.numpy(), lanes=el, layout=1)
        _ = float(yh[-1])
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
    e2e = {"value": world * el * en * 16 * 3 / (e2e_ms * 1e-3) / 1e9, "unit": "GSa/s", "h2d_bytes_per_step": 4 * el * en * 16,
           "d2h_bytes_per_step": 4 * el * en * 16, "api": "idsp_chain_f32_host (pinned host buffers, dec -> int -> biquad in one round trip)",
           "lanes": el, "samples_per_lane": en * 16}
    del xh, yh
    if rank != 0:
        return None
    peak, peak_src = peak_hbm()
    best = max(points, key=lambda p: p["GSa/s"])
    return {
        "metric": "GSa/s, f32 HbfDec(/16)->HbfInt(x16)->Biquad DF1 chain, lane sweep", "value": best["GSa/s"], "unit": "GSa/s",
        "n_gpus": world, "steps": 3, "warmup": 2, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[4]: HbfDec->HbfInt->Biquad chain f32, lanes 2^10..2^24, ~2^30 samples per point, lane-major"},
        "sweep": points, "e2e": e2e,
        "roofline": {"bound": "hbm", "achieved": best["GB/s"], "peak": peak, "unit": "GB/s", "frac": best["GB/s"] / peak,
                     "traffic": None, "peak_source": peak_src},
        "parity_check": "2^10-, 2^16- and 2^20-lane points == oracle on 24+ lanes strided over the lane range x up to 8192 samples",
    }


def run_plumbing(args, rank, world, local):
    """configs[0]: 1-lane f32 iir::Biquad DF2T lowpass over 1e6 white-noise samples.  The reference
    crate cannot be built here, so its scalar loop is the C port on ONE host core; the same stream
    then goes through the C ABI on the GPU (one lane = one thread: pure latency, plumbing only) and
    must produce the same bits."""
    import hashlib

    import torch

    import oracle as O
    from idsp_b200 import Biquad, DirectForm2Transposed, Filter, Lanes

    n = 1_000_000
    bq = Biquad.from_ba6(Filter().critical_frequency(0.01).lowpass(), "f32")
    x = np.random.default_rng(1).standard_normal(n).astype(np.float32)
    st = np.zeros(2, np.float32)
    O.biquad_df2t("f32", bq.ba, None, st.copy(), x[:1000])
    t0 = time.perf_counter()
    reps = max(1, args.steps // 10)
    for _ in range(reps):
        s = st.copy()
        want = O.biquad_df2t("f32", bq.ba, None, s, x)
    cpu_s = (time.perf_counter() - t0) / reps
    dev = f"cuda:{local}"
    xd = torch.from_numpy(x).to(dev)
    yd = torch.empty_like(xd)
    sd = DirectForm2Transposed.default("f32", 1, dev)
    Lanes(bq).block(sd, xd, yd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sd = DirectForm2Transposed.default("f32", 1, dev)
    e0.record()
    Lanes(bq).block(sd, xd, yd)
    e1.record()
    torch.cuda.synchronize()
    got = yd.cpu().numpy()
    same = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
    if not same:
        raise SystemExit("bench plumbing: GPU output differs from the CPU port")
    if rank != 0:
        return None
    return {
        "metric": "MSa/s, 1-lane f32 iir::Biquad DF2T lowpass, 1e6 white-noise samples", "value": n / cpu_s / 1e6, "unit": "MSa/s",
        "n_gpus": world, "steps": reps, "warmup": 1, "ms_per_step": cpu_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[0]: 1-lane f32 Biquad DF2T lowpass f0=0.01, 1e6 N(0,1) samples (seed 1), host CPU, 1 core"},
        "cpu_baseline": {"value": n / cpu_s / 1e6, "unit": "MSa/s", "cores": 1, "kind": "port", "sample": "the whole 1e6-sample stream"},
        "gpu_one_lane": {"value": n / (e0.elapsed_time(e1) * 1e-3) / 1e6, "unit": "MSa/s",
                         "note": "one lane = one GPU thread: latency bound by construction; parity only"},
        "checksum_sha256": hashlib.sha256(want.tobytes()).hexdigest(), "gpu_bits_equal_cpu": same,
    }


def run_biquad(args, rank, world, local):
    import torch

    import oracle as O
    from idsp_b200 import DirectForm1, Lanes
    from idsp_b200.engine import default_context

    dev = f"cuda:{local}"
    ctx = default_context(local)
    bq = biquad_coeffs()
    cfg = Lanes(bq)
    lanes, frames = BIQUAD_LANES, args.frames
    n = lanes * frames
    nring = args.ring
    gen = torch.Generator(device=dev)
    gen.manual_seed(2 + rank)
    xin = [torch.randint(-(1 << 28), 1 << 28, (n,), dtype=torch.int32, device=dev, generator=gen) for _ in range(nring)]
    yout = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2)]
    st = DirectForm1.default("i32", lanes, dev)
    steps = args.steps

    layout = args.layout

    def step(i):
        cfg.block(st, xin[i % nring], yout[i % 2], layout)

    # parity gate: first step against the oracle on lanes strided over the whole launch (every CTA's box is
    # sampled somewhere, plus the first and last 8 lanes), all frames, and the filter state after the step
    step(0)
    torch.cuda.synchronize()
    idx_np = strided_lanes(lanes, 80)
    idx = torch.from_numpy(idx_np).to(dev)
    sub = int(idx.numel())
    if layout == 0:
        xs = xin[0].view(frames, lanes).index_select(1, idx).contiguous().cpu().numpy()
        got = yout[0].view(frames, lanes).index_select(1, idx).contiguous().cpu().numpy().reshape(-1)
    else:
        xs = xin[0].view(lanes, frames).index_select(0, idx).contiguous().cpu().numpy()
        got = yout[0].view(lanes, frames).index_select(0, idx).contiguous().cpu().numpy().reshape(-1)
    so = np.zeros((4, sub), np.int32)
    want = O.biquad_lanes("df1", "i32", bq.ba, F_BITS, None, so, xs.reshape(-1), sub, layout, nthreads=host_threads())
    if not np.array_equal(got, want) or not np.array_equal(st.numpy()[:, idx_np], so):
        raise SystemExit("bench: GPU output differs from the oracle -- refusing to report a number")
    kernel_family = ctx.last_kernel

    for i in range(args.warmup):
        step(i + 1)
    sampler = ClockSampler(local)
    barrier(world)
    l0 = ctx.launches
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    launches = ctx.launches - l0
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    value = world * n * steps / (ms * 1e-3) / 1e9
    if args.profile or args.full:
        keep_busy_until_sampled(sampler, step)
        sus_ms, sus_n = None, 0
    else:
        # the K timed steps are a ~25 ms burst at boost clocks: the same step for >= 2 s shows the rate a long
        # job sustains under the power cap (the clock sampler keeps running through it)
        sus_ms, sus_n = sustained_loop(step, ms / steps, world, dev, args.sustained_seconds)
    clocks = sampler.stop()
    barrier(world)

    if args.profile:
        if rank == 0:
            return {"profile_run": True, "value": value, "ms_per_step": ms / steps, "gpu_launches": int(launches)}
        return None
    # e2e: C ABI host entry point, pinned host buffers, H2D + D2H inside the timed region
    ef = args.e2e_frames
    xh = torch.randint(-(1 << 28), 1 << 28, (ef * lanes,), dtype=torch.int32).pin_memory()
    yh = torch.empty(ef * lanes, dtype=torch.int32).pin_memory()
    sth = DirectForm1.default("i32", lanes, None)
    xa, ya = xh.numpy(), yh.numpy()
    cfg.block(sth, xa, ya, layout)  # warm-up (allocates the staging ring)
    barrier(world)
    esteps = max(1, min(steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(esteps):
        cfg.block(sth, xa, ya, layout)
        _ = int(ya[-1])  # device -> host result is read
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
    e2e = world * ef * lanes * esteps / (e2e_ms * 1e-3) / 1e9
    h2d_gbs, d2h_gbs = pcie_peak(dev, world=world)

    del xin, yout, xh, yh
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = peak_hbm()
    per_launch_bytes = 8.0 * n  # 4 B read + 4 B written per sample; state/coeff traffic ~0
    achieved = per_launch_bytes * launches / (ms * 1e-3) / 1e9 if launches else 0.0
    cpu_v, cpu_dt, cpu_sample = cpu_calibrated("biquad", host_threads(), args.cpu_seconds) if world == 1 else (None, None, "measured at N=1 only")
    cpu1_v = cpu_calibrated("biquad", 1, max(2.0, args.cpu_seconds / 4))[0] if world == 1 else None
    line = {
        "metric": metric_name("biquad"), "value": value, "unit": "GSa/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32", "data": "synthetic", "config": biquad_config(args, frames),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic_from_profiles("biquad_df1_i32_fm_bytes_per_launch"),
                     "peak_source": peak_src, "kernel": "tma_lanes_kernel<Df1Op<int,false,1>> (frame-major: 256-lane x 8-frame TMA boxes)" if layout == 0 else "tma_lanes_kernel<Df1Op<int,false,1>> (lane-major: swizzled 16-frame x 32-lane TMA boxes)",
                     "kernel_family": kernel_family,
                     "algorithmic_bytes_per_launch": per_launch_bytes,
                     "sustained": None if sus_ms is None else {
                         "what": f"the same step repeated {sus_n} times (>= {args.sustained_seconds} s), power-capped clocks",
                         "value_GSa/s": world * n / (sus_ms * 1e-3) / 1e9, "achieved": per_launch_bytes / (sus_ms * 1e-3) / 1e9,
                         "frac": per_launch_bytes / (sus_ms * 1e-3) / 1e9 / peak, "steps": sus_n}},
        "cpu_baseline": {"value": cpu_v, "unit": "GSa/s", "cores": host_threads(), "kind": "port",
                         "sample": cpu_sample, "seconds": cpu_dt, "one_core": {"value": cpu1_v, "unit": "GSa/s", "cores": 1},
                         "build": O.build_flags() + " -ffp-contract=off -fopenmp (built on this host)"},
        "e2e": {"value": e2e, "unit": "GSa/s", "h2d_bytes_per_step": 4 * ef * lanes, "d2h_bytes_per_step": 4 * ef * lanes,
                "frames_per_step": ef, "steps": esteps, "api": "idsp_biquad_df1_i32_host (pinned host buffers)",
                "pcie_gbs": {"h2d": h2d_gbs, "d2h": d2h_gbs,
                             "how": "512 MiB pinned copies, both directions at once, all ranks at the same time (barrier), slowest rank"},
                "bound": "PCIe: 4 B in + 4 B out per sample; the yardstick is the mean of the two directions (under contention they share one limit: the host's memory system)",
                "frac_of_pcie": (e2e / world) * 4.0 / (0.5 * (h2d_gbs + d2h_gbs))},
        "gpu_launches": int(launches), "clocks": clocks,
        "parity_check": f"first step == oracle on {sub} lanes strided over all {lanes} lanes (every TMA box family, first / last 8) x all frames + state",
    }
    if layout == 1:
        line["config"]["workload"] = line["config"]["workload"].replace("frame-major", "lane-major")
    if args.full:
        line["full_job"] = {"frames": steps * frames, "seconds": ms * 1e-3}
    return line


def verify_full_job(args, rank, local):
    """--full: replay the whole 1e7-frame job on a lane subset with the CPU oracle (same cyclic ring
    of input blocks, state carried from the very first frame) and compare the FINAL filter state
    and the last output block bit for bit."""
    import torch

    import oracle as O
    from idsp_b200 import DirectForm1, Lanes

    dev = f"cuda:{local}"
    bq = biquad_coeffs()
    cfg = Lanes(bq)
    lanes, frames, nring, sub = BIQUAD_LANES, args.frames, args.ring, 64
    n = lanes * frames
    gen = torch.Generator(device=dev)
    gen.manual_seed(2 + rank)
    xin = [torch.randint(-(1 << 28), 1 << 28, (n,), dtype=torch.int32, device=dev, generator=gen) for _ in range(nring)]
    y = torch.empty(n, dtype=torch.int32, device=dev)
    st = DirectForm1.default("i32", lanes, dev)
    steps = (TOTAL_FRAMES + frames - 1) // frames
    for i in range(steps):
        cfg.block(st, xin[i % nring], y, 0)
    torch.cuda.synchronize()
    xs = [x.view(frames, lanes)[:, :sub].contiguous().cpu().numpy().reshape(-1) for x in xin]
    so = np.zeros((4, sub), np.int32)
    t0 = time.perf_counter()
    for i in range(steps):
        want = O.biquad_lanes("df1", "i32", bq.ba, F_BITS, None, so, xs[i % nring], sub, 0, nthreads=host_threads())
    cpu_s = time.perf_counter() - t0
    got = y.view(frames, lanes)[:, :sub].contiguous().cpu().numpy().reshape(-1)
    ok = bool(np.array_equal(got, want) and np.array_equal(st.numpy()[:, :sub], so))
    return {"lanes_checked": sub, "frames": steps * frames, "final_state_and_last_block_bit_exact": ok,
            "cpu_replay_seconds": cpu_s}


def run_hbf(args, rank, world, local):
    import torch

    import oracle as O
    from idsp_b200 import HbfDecCascade, Lanes
    from idsp_b200.engine import default_context
    from idsp_b200.hbf import HbfDec16

    dev = f"cuda:{local}"
    ctx = default_context(local)
    cfg = Lanes(HbfDecCascade(4))
    slices = args.hbf_slices
    lanes_s = HBF_LANES // slices
    n_out = HBF_INPUTS // 16
    n_in = lanes_s * HBF_INPUTS
    nring = min(slices, args.ring)
    gen = torch.Generator(device=dev)
    gen.manual_seed(3 + rank)
    xin = []
    for _ in range(nring):
        t = torch.empty(n_in, dtype=torch.float32, device=dev)
        t.uniform_(-1, 1, generator=gen)
        xin.append(t)
    yout = [torch.empty(lanes_s * n_out, dtype=torch.float32, device=dev) for _ in range(2)]
    states = [HbfDec16(lanes_s, dev) for _ in range(2)]

    hl = args.hbf_layout  # x: lane-major [lane][65536] or frame-major [4096 frames][lane][16]

    def step(i):
        states[i % 2].words.zero_()
        cfg.block(states[i % 2], xin[i % nring], yout[i % 2], hl)

    step(0)
    torch.cuda.synchronize()
    idx = torch.from_numpy(strided_lanes(lanes_s, 40)).to(dev)
    sub = int(idx.numel())
    if hl == 1:
        xs = xin[0].view(lanes_s, HBF_INPUTS).index_select(0, idx).contiguous().cpu().numpy().reshape(-1)
        got = yout[0].view(lanes_s, n_out).index_select(0, idx).contiguous().cpu().numpy().reshape(-1)
    else:
        xs = xin[0].view(n_out, lanes_s, 16).index_select(1, idx).contiguous().cpu().numpy().reshape(-1)
        got = yout[0].view(n_out, lanes_s).index_select(1, idx).contiguous().cpu().numpy().reshape(-1)
    so = np.zeros((O.hbf_dec_state_words(4), sub), np.float32)
    want = O.hbf_dec_cascade_lanes(4, so, xs, sub, hl, nthreads=host_threads())
    if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
        raise SystemExit("bench: GPU HBF output differs from the oracle -- refusing to report a number")
    for i in range(args.warmup):
        step(i + 1)
    sampler = ClockSampler(local)
    barrier(world)
    l0 = ctx.launches
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    launches = ctx.launches - l0
    keep_busy_until_sampled(sampler, step)
    clocks = sampler.stop()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    value = world * n_in * args.steps / (ms * 1e-3) / 1e9

    if args.profile:
        if rank == 0:
            return {"profile_run": True, "value": value, "ms_per_step": ms / args.steps, "gpu_launches": int(launches)}
        return None
    if getattr(args, "resident_only", False):  # the second layout of the default run: resident leg + roofline only
        del xin, yout
        torch.cuda.empty_cache()
        if rank != 0:
            return None
        peak, peak_src = peak_hbm()
        per_launch_bytes = 4.25 * n_in
        achieved = per_launch_bytes * launches / (ms * 1e-3) / 1e9 if launches else 0.0
        cfg_h = hbf_config(args)
        if hl == 0:
            cfg_h["workload"] = cfg_h["workload"].replace("lane-major", "frame-major ([[f32;16]; lanes] per frame)")
        return {"metric": metric_name("hbf"), "value": value, "unit": "GSa/s", "n_gpus": world, "steps": args.steps,
                "ms_per_step": ms / args.steps, "dtype": "f32", "config": cfg_h, "kernel": ctx.last_kernel,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic_from_profiles("hbf_dec16_f32_fm_bytes_per_launch") if hl == 0 else None,
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch_bytes},
                "gpu_launches": int(launches), "clocks": clocks,
                "parity_check": f"first step == oracle on {sub} lanes strided over the slice (first / last 8 included) x all samples"}
    el = 8192
    xh = torch.empty(el * HBF_INPUTS, dtype=torch.float32).uniform_(-1, 1).pin_memory()
    yh = torch.empty(el * n_out, dtype=torch.float32).pin_memory()
    sth = HbfDec16(el, None)
    cfg.block(sth, xh.numpy(), yh.numpy(), 1)
    barrier(world)
    esteps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(esteps):
        cfg.block(sth, xh.numpy(), yh.numpy(), 1)
        _ = float(yh[-1])
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
    e2e = world * el * HBF_INPUTS * esteps / (e2e_ms * 1e-3) / 1e9
    h2d_gbs, d2h_gbs = pcie_peak(dev, world=world, bidir=False)
    del xin, yout, xh, yh
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = peak_hbm()
    per_launch_bytes = 4.25 * n_in
    achieved = per_launch_bytes * launches / (ms * 1e-3) / 1e9 if launches else 0.0
    cpu_v, cpu_dt, cpu_sample = cpu_calibrated("hbf", host_threads(), args.cpu_seconds) if world == 1 else (None, None, "measured at N=1 only")
    cfg_h = hbf_config(args)
    if hl == 0:
        cfg_h["workload"] = cfg_h["workload"].replace("lane-major", "frame-major ([[f32;16]; lanes] per frame)")
    line = {
        "metric": metric_name("hbf"), "value": value, "unit": "GSa/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_h,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic_from_profiles("hbf_dec16_f32_lm_bytes_per_launch") if hl == 1 else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": per_launch_bytes},
        "cpu_baseline": {"value": cpu_v, "unit": "GSa/s", "cores": host_threads(), "kind": "port",
                         "sample": cpu_sample, "seconds": cpu_dt},
        "e2e": {"value": e2e, "unit": "GSa/s", "h2d_bytes_per_step": 4 * el * HBF_INPUTS, "d2h_bytes_per_step": 4 * el * n_out,
                "steps": esteps, "api": "idsp_hbf_dec_cascade_f32_host (pinned host buffers)",
                "pcie_gbs": {"h2d": h2d_gbs, "d2h": None,
                             "how": "512 MiB pinned host -> device copies with 1/16 of that going back at the same time, all ranks at the same time (barrier), slowest rank"},
                "bound": "PCIe: 4 B in + 0.25 B out per sample", "frac_of_pcie": (e2e / world) * 4.0 / h2d_gbs},
        "gpu_launches": int(launches), "clocks": clocks,
        "parity_check": f"first step == oracle on {sub} lanes strided over the slice (first / last 8 included) x all samples",
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="biquad", choices=["biquad", "hbf", "lockin", "lockin_sharded", "chain", "plumbing"])
    ap.add_argument("--frames", type=int, default=16384, help="frames per step (biquad)")
    ap.add_argument("--ring", type=int, default=3, help="distinct resident input blocks")
    ap.add_argument("--full", action="store_true", help="run the whole 1e7-frame job (biquad)")
    ap.add_argument("--hbf-slices", type=int, default=8)
    ap.add_argument("--e2e-frames", type=int, default=2048)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the sustained-rate loop after the timed steps")
    ap.add_argument("--sharded-frames", type=int, default=4096, help="frames per pass of the sharded lock-in leg (root-resident)")
    ap.add_argument("--layout", type=int, default=None, choices=[0, 1],
                    help="0 frame-major, 1 lane-major (default: biquad 0, hbf 1)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary (hbf) measurement of the default run")
    ap.add_argument("--profile", action="store_true", help="profiling run: skip the CPU baseline and e2e legs")
    args = ap.parse_args()
    args.hbf_layout = 1 if args.layout is None else args.layout
    args.layout = 0 if args.layout is None else args.layout
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.full:
        args.steps = (TOTAL_FRAMES + args.frames - 1) // args.frames
    if args.impl == "reference":
        run_reference(args)
        return
    rank, world, local = dist_setup(args.gpus)
    if args.workload == "biquad":
        line = run_biquad(args, rank, world, local)
        if args.full and line is not None and not args.profile:
            line["full_job"]["verify"] = verify_full_job(args, rank, local)
            if not line["full_job"]["verify"]["final_state_and_last_block_bit_exact"]:
                raise SystemExit("bench --full: final state differs from the oracle")
        if not (args.no_extra or args.profile or args.full):
            # secondary headline (BASELINE configs[2]) measured in the same run, same contract
            import copy

            a2 = copy.copy(args)
            a2.steps = min(args.steps, 24)
            extra = run_hbf(a2, rank, world, local)
            if line is not None and extra is not None:
                line["extra"] = {"hbf_dec16_f32": {k: extra[k] for k in ("metric", "value", "unit", "steps", "ms_per_step", "dtype", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "parity_check")}}
            # BASELINE configs[3] (resident per GPU, and as the sharded data plane over NVLink) and configs[4] in
            # the same run at every N; a failure here must not cost the headline line, so it is recorded
            # instead of raised (every rank takes the same path: the legs contain collectives)
            import torch

            if line is not None:
                line.setdefault("extra", {})
            def run_hbf_fm(a, rank, world, local):  # configs[2] on the reference's own frame format [[f32; 16]; lanes]
                a.hbf_layout, a.resident_only = 0, True  # resident leg only (the host / CPU legs are in hbf_dec16_f32)
                return run_hbf(a, rank, world, local)

            for name, fn, keys in (("hbf_dec16_f32_frame_major", run_hbf_fm, None),
                                   ("lockin_i32", run_lockin, ("metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "dtype", "config", "roofline", "gpu_launches", "parity_check")),
                                   ("lockin_sharded", run_lockin_sharded, None),
                                   ("chain_f32", run_chain, ("metric", "value", "unit", "n_gpus", "dtype", "config", "sweep", "roofline", "e2e", "parity_check"))):
                try:
                    torch.cuda.empty_cache()
                    a3 = copy.copy(args)
                    a3.steps = min(args.steps, 24)
                    res = fn(a3, rank, world, local)
                    if line is not None and res is not None:
                        line["extra"][name] = res if keys is None else {k: res[k] for k in keys if k in res}
                except (Exception, SystemExit) as e:  # noqa: BLE001
                    if world > 1:
                        raise  # a rank that left a collective leg cannot rejoin: fail loudly instead of hanging
                    if line is not None:
                        line["extra"][name] = {"error": f"{type(e).__name__}: {e}"[:300]}
    elif args.workload == "hbf":
        line = run_hbf(args, rank, world, local)
    elif args.workload == "lockin":
        line = run_lockin(args, rank, world, local)
    elif args.workload == "lockin_sharded":
        line = run_lockin_sharded(args, rank, world, local)
    elif args.workload == "plumbing":
        line = run_plumbing(args, rank, world, local)
    else:
        line = run_chain(args, rank, world, local)
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
