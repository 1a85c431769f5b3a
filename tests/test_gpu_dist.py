"""N > 1 on real GPUs (BASELINE configs[3] shape): rank 0 holds the samples of all lanes,
NCCL scatters contiguous lane blocks over NVLink, every rank runs the fused lock-in on its own lanes
(no collective inside the computation), NCCL gathers the Complex<i32> outputs, and rank 0 compares them with
the CPU oracle bit for bit.  Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu -s`);
skipped on a single-GPU box.  The gloo / CPU version of the same plumbing is tests/test_dist_gloo.py."""
import os
import socket
import time

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

K = [1048576, -94906265]  # Lowpass<2>, SURVEY 8(d) cfg 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frames, lanes, layout, ok, out_path):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        from idsp_b200 import Accu, Lockin, LockinState, Lowpass
        from idsp_b200.dist import gather_lanes, lane_block, scatter_lanes

        x = a0 = step = None
        if rank == 0:
            rng = np.random.default_rng(4)
            xn = rng.integers(-(1 << 30), 1 << 30, frames * lanes).astype(np.int32)
            a0n = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
            stn = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
            x, a0, step = (torch.from_numpy(v).to(dev) for v in (xn, a0n, stn))
        lo, hi = lane_block(rank, world, lanes)
        # warm the NCCL point-to-point channels, then time the edges on the device
        gather_lanes(scatter_lanes(a0, 1, lanes, 0, dtype=torch.int32, device=dev), 1, lanes, 0)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        xs = scatter_lanes(x, frames, lanes, layout, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        ts = time.perf_counter()
        a0s = scatter_lanes(a0, 1, lanes, 0, dtype=torch.int32, device=dev)
        sts = scatter_lanes(step, 1, lanes, 0, dtype=torch.int32, device=dev)
        iq = torch.empty(2 * xs.numel(), dtype=torch.int32, device=dev)
        Lockin(Lowpass(K)).block(LockinState.default(2, hi - lo, dev), Accu(a0s.clone(), sts), xs, iq, layout)  # warm-up
        torch.cuda.synchronize()
        dist.barrier()
        t1 = time.perf_counter()
        st = LockinState.default(2, hi - lo, dev)
        acc = Accu(a0s, sts)
        Lockin(Lowpass(K)).block(st, acc, xs, iq, layout)
        torch.cuda.synchronize()
        dist.barrier()
        t2 = time.perf_counter()
        full = gather_lanes(iq, frames, lanes, layout, width=2)
        torch.cuda.synchronize()
        dist.barrier()
        t3 = time.perf_counter()
        if rank == 0:
            import oracle as O

            O.build()
            ao = a0n.copy()
            so = np.zeros((4, lanes), np.int64)
            want = O.lockin_lanes(K, ao, stn, so, xn, lanes, layout, nthreads=8)
            got = full.cpu().numpy()
            assert np.array_equal(got, want), "gathered lock-in output differs from the oracle"
            n = frames * lanes
            line = (f"world={world} layout={layout} lanes={lanes} frames={frames}: scatter {4 * n / (ts - t0) / 1e9:.1f} GB/s, "
                    f"lock-in {n / (t2 - t1) / 1e9:.1f} GSa/s, gather {8 * n / (t3 - t2) / 1e9:.1f} GB/s, bit-exact")
            print(line, flush=True)
            if out_path:
                with open(out_path, "a") as f:
                    f.write(line + "\n")
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("layout,frames,lanes_per_gpu", [(0, 512, 8192), (1, 512, 8192), (0, 2048, 131072)])
def test_scatter_lockin_gather_nccl(layout, frames, lanes_per_gpu):
    """the last case has BASELINE configs[3]'s 131 072 lanes per GPU (the timings it prints are meaningful)"""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    lanes = lanes_per_gpu * world + 40  # ragged last block
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    port = _free_port()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    out_path = os.path.join(out, "dist_nccl.log") if os.path.isdir(out) else None
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, lanes, layout, ok, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert list(ok) == [1] * world


def _worker_peer(rank, world, port, frames, lanes, ok, out_path):
    """lane-major lock-in whose output goes straight into rank 0's buffer (CUDA IPC peer stores from the
    kernel epilogue) vs the same kernel followed by an NCCL gather"""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        from idsp_b200 import Accu, Lockin, LockinState, Lowpass
        from idsp_b200.dist import PeerBuffer, gather_lanes, lane_block, scatter_lanes

        layout = 1
        x = a0 = step = None
        if rank == 0:
            rng = np.random.default_rng(5)
            xn = rng.integers(-(1 << 30), 1 << 30, frames * lanes).astype(np.int32)
            a0n = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
            stn = rng.integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
            x, a0, step = (torch.from_numpy(v).to(dev) for v in (xn, a0n, stn))
        lo, hi = lane_block(rank, world, lanes)
        xs = scatter_lanes(x, frames, lanes, layout, dtype=torch.int32, device=dev)
        a0s = scatter_lanes(a0, 1, lanes, 0, dtype=torch.int32, device=dev)
        sts = scatter_lanes(step, 1, lanes, 0, dtype=torch.int32, device=dev)
        buf = PeerBuffer(2 * frames * lanes, torch.int32, rank, owner=0)
        mine = buf.view(2 * frames * lo, 2 * frames * (hi - lo))
        local = torch.empty(2 * xs.numel(), dtype=torch.int32, device=dev)
        lock = Lockin(Lowpass(K))

        def run(out):
            lock.block(LockinState.default(2, hi - lo, dev), Accu(a0s.clone(), sts), xs, out, layout)

        run(local)
        gather_lanes(local, frames, lanes, layout, width=2)  # warm-up of both paths
        run(mine)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        run(local)
        full = gather_lanes(local, frames, lanes, layout, width=2)
        torch.cuda.synchronize()
        dist.barrier()
        t1 = time.perf_counter()
        run(mine)
        torch.cuda.synchronize()
        dist.barrier()
        t2 = time.perf_counter()
        if rank == 0:
            import oracle as O

            O.build()
            ao = a0n.copy()
            so = np.zeros((4, lanes), np.int64)
            want = O.lockin_lanes(K, ao, stn, so, xn, lanes, layout, nthreads=8)
            assert np.array_equal(full.cpu().numpy(), want), "NCCL-gathered output differs from the oracle"
            assert np.array_equal(buf.tensor().cpu().numpy(), want), "peer-stored output differs from the oracle"
            n = frames * lanes
            line = (f"world={world} lane-major lanes={lanes} frames={frames}: kernel + NCCL gather {n / (t1 - t0) / 1e9:.1f} GSa/s, "
                    f"kernel storing into rank 0's buffer over NVLink {n / (t2 - t1) / 1e9:.1f} GSa/s, both bit-exact")
            print(line, flush=True)
            if out_path:
                with open(out_path, "a") as f:
                    f.write(line + "\n")
        buf.close()
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_lockin_stores_into_root_buffer_over_nvlink():
    """SURVEY 8(e) fused form: the result tiles leave the kernel epilogue for the root's buffer"""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    frames, lanes = 2048, 131072 * world
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    port = _free_port()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    out_path = os.path.join(out, "dist_nccl.log") if os.path.isdir(out) else None
    procs = [ctx.Process(target=_worker_peer, args=(r, world, port, frames, lanes, ok, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert list(ok) == [1] * world
