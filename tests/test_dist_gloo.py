"""N > 1 plumbing on CPU: world_size-2 gloo processes exercise the lane partition and the
scatter / gather of lane blocks for both layouts (the compute itself needs a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from idsp_b200.dist import all_blocks, gather_lanes, lane_block, scatter_lanes, shard_flat, shard_state


def test_lane_blocks_cover_and_balance():
    for lanes in (1, 31, 32, 33, 1000, 65536, 1048576 + 5):
        for world in (1, 2, 3, 4, 8):
            b = all_blocks(world, lanes)
            assert b[0][0] == 0 and b[-1][1] == lanes
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 32 or lanes < 32 * world
            assert all(lo % 32 == 0 for lo, _ in b)


def test_shard_flat_layouts():
    frames, lanes, w = 5, 7, 2
    x = torch.arange(frames * lanes * w).view(frames, lanes, w)
    fm = x.reshape(-1)
    lm = x.permute(1, 0, 2).reshape(-1)
    assert torch.equal(shard_flat(fm, frames, lanes, 2, 5, 0, w).view(frames, 3, w), x[:, 2:5])
    assert torch.equal(shard_flat(lm, frames, lanes, 2, 5, 1, w).view(3, frames, w), x[:, 2:5].permute(1, 0, 2))
    st = torch.arange(4 * lanes).view(4, lanes)
    assert torch.equal(shard_state(st, 2, 5), st[:, 2:5])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frames, lanes, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for layout in (0, 1):
            for width in (1, 2):
                full = None
                if rank == 0:
                    g = torch.Generator().manual_seed(layout * 2 + width)
                    full = torch.randint(-1000, 1000, (frames * lanes * width,), dtype=torch.int32, generator=g)
                part = scatter_lanes(full, frames, lanes, layout, width, dtype=torch.int32, device="cpu")
                lo, hi = lane_block(rank, world, lanes)
                assert part.numel() == frames * (hi - lo) * width
                # "compute": a per-lane independent op (negate) stands in for the filter
                back = gather_lanes(-part, frames, lanes, layout, width)
                if rank == 0:
                    assert torch.equal(back, -full)
        # max-over-ranks reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lanes", [70, 64])
def test_scatter_gather_world2(lanes):
    world = 2
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 6, lanes, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world
