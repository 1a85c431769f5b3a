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


# ------------------------------------------------------------------ the C ABI's own communicator (raw NCCL, round 2)
def _worker_cabi(rank, world, port, ok, out_path):
    """idsp_b200_comm_init / idsp_scatter_lanes / idsp_gather_lanes / idsp_broadcast: every dtype width, both
    layouts, ragged lane counts (blocks of different sizes, an empty block), then the sharded lock-in through it.
    torch.distributed only carries the 128-byte NCCL id (gloo: the data plane must not depend on torch's NCCL)."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from idsp_b200 import Accu, Lockin, LockinState, Lowpass, _lib
        from idsp_b200.dist import Comm, all_blocks

        assert _lib.lib().idsp_b200_nccl_version() >= 20700
        comm = Comm(rank)
        rng = np.random.default_rng(11)  # same stream on every rank: everyone can compute the expected block
        for dtype, width in ((torch.int32, 1), (torch.int32, 2), (torch.float32, 16), (torch.int64, 1), (torch.int8, 1)):
            for layout in (0, 1):
                for frames, lanes in ((7, 40 * world + 5), (64, 32), (33, 1000), (1, 32 * world)):
                    n = frames * lanes * width
                    src = torch.from_numpy(rng.integers(-100, 100, n).astype(np.int64)).to(dtype)
                    lo, hi = all_blocks(world, lanes)[rank]
                    assert (lo, hi) == comm.lane_block(lanes)
                    full = src.to(dev) if rank == 0 else None
                    part = comm.scatter_lanes(full, frames, lanes, layout, width=width, dtype=dtype)
                    v = src.view(frames, lanes, width) if layout == 0 else src.view(lanes, frames, width).permute(1, 0, 2)
                    want = v[:, lo:hi] if layout == 0 else v[:, lo:hi].permute(1, 0, 2)
                    torch.cuda.synchronize()
                    assert torch.equal(part.cpu(), want.contiguous().view(-1)), (dtype, width, layout, frames, lanes)
                    back = comm.gather_lanes(part + 1 if dtype != torch.float32 else part + 1.0, None, frames, lanes, layout, width=width)
                    torch.cuda.synchronize()
                    if rank == 0:
                        assert torch.equal(back.cpu(), src + 1), (dtype, width, layout, frames, lanes)
        b = torch.arange(5, dtype=torch.int32, device=dev) * (1 if rank == 0 else 0)
        comm.broadcast(b)
        torch.cuda.synchronize()
        assert b.cpu().tolist() == [0, 1, 2, 3, 4]
        # sharded lock-in through the C ABI edges, frame-major (packed blocks) and lane-major (in place)
        for layout in (0, 1):
            frames, lanes = 256, 4096 * world + 96
            xn = np.random.default_rng(4).integers(-(1 << 30), 1 << 30, frames * lanes).astype(np.int32)
            stn = np.random.default_rng(5).integers(-(1 << 31), 1 << 31, lanes).astype(np.int32)
            lo, hi = comm.lane_block(lanes)
            x = torch.from_numpy(xn).to(dev) if rank == 0 else None
            step = torch.from_numpy(stn).to(dev) if rank == 0 else None
            xs = comm.scatter_lanes(x, frames, lanes, layout)
            sts = comm.scatter_lanes(step, 1, lanes, 0)
            iq = torch.empty(2 * xs.numel(), dtype=torch.int32, device=dev)
            Lockin(Lowpass(K)).block(LockinState.default(2, hi - lo, dev), Accu(torch.zeros(hi - lo, dtype=torch.int32, device=dev), sts), xs, iq, layout)
            t0 = time.perf_counter()
            full = comm.gather_lanes(iq, None, frames, lanes, layout, width=2)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            if rank == 0:
                import oracle as O

                O.build()
                want = O.lockin_lanes(K, np.zeros(lanes, np.int32), stn, np.zeros((4, lanes), np.int64), xn, lanes, layout, nthreads=8)
                assert np.array_equal(full.cpu().numpy(), want), "C-ABI gathered lock-in output differs from the oracle"
                line = f"world={world} C-ABI comm (NCCL {_lib.lib().idsp_b200_nccl_version()}): layout={layout} lanes={lanes} frames={frames} scatter -> lock-in -> gather bit-exact ({8 * frames * lanes / (t1 - t0) / 1e9:.1f} GB/s gather incl. launch)"
                print(line, flush=True)
                if out_path:
                    with open(out_path, "a") as f:
                        f.write(line + "\n")
        comm.close()
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_c_abi_comm_scatter_gather_broadcast():
    """SURVEY 8(b) `idsp_b200_comm_init / idsp_scatter_lanes / idsp_gather_lanes` under dsp-process/src/split.rs:272-277"""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    port = _free_port()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    out_path = os.path.join(out, "dist_nccl.log") if os.path.isdir(out) else None
    procs = [ctx.Process(target=_worker_cabi, args=(r, world, port, ok, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert list(ok) == [1] * world


def test_c_abi_comm_single_rank():
    """nranks == 1 never touches NCCL: scatter / gather are the (strided) local copies"""
    from idsp_b200.dist import Comm

    comm = Comm(0)
    assert (comm.rank, comm.world) == (0, 1)
    for layout in (0, 1):
        src = torch.arange(6 * 50 * 2, dtype=torch.int32, device="cuda:0")
        part = comm.scatter_lanes(src, 6, 50, layout, width=2)
        torch.cuda.synchronize()
        assert torch.equal(part, src)
        full = comm.gather_lanes(part, None, 6, 50, layout, width=2)
        torch.cuda.synchronize()
        assert torch.equal(full, src)
    comm.close()
