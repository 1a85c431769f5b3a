"""Lane sharding across the GPUs of one node (one process per GPU, torch.distributed).

Lanes never interact (dsp-process/src/compose.rs:472-475), so the path shards by
independent units: rank r owns the contiguous lane block [lo, hi) and the matching
slice of the SoA state; coefficients are tiny and replicated.  There is no collective
inside the computation.  NCCL (or gloo in the CPU tests) is used only at the edges:
`scatter_lanes` hands lane blocks of a root-resident buffer to the ranks,
`gather_lanes` collects outputs, both for either layout of dsp-process/src/view.rs.
`PeerBuffer` is the fused form of the gather: the root's result buffer is mapped into every
process (CUDA IPC), and each rank passes its lane block of it as the output of its kernel, whose
stores then travel over NVLink / NVSwitch from the kernel's own epilogue.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

import ctypes as C

from . import _lib
from .engine import FRAME_MAJOR, LANE_MAJOR, PeerView, default_context


def lane_block(rank: int, world: int, lanes: int, align: int = 32) -> Tuple[int, int]:
    """Contiguous lane block of `rank`: blocks are multiples of `align` lanes (a warp) except
    the last, cover [0, lanes) exactly and differ by at most one unit."""
    units = (lanes + align - 1) // align
    lo_u = units * rank // world
    hi_u = units * (rank + 1) // world
    return min(lo_u * align, lanes), min(hi_u * align, lanes)


def all_blocks(world: int, lanes: int, align: int = 32) -> List[Tuple[int, int]]:
    return [lane_block(r, world, lanes, align) for r in range(world)]


def _as_tlw(flat: torch.Tensor, frames: int, lanes: int, width: int, layout: int) -> torch.Tensor:
    """view a flat buffer as [frames, lanes, width] (a permuted view for lane-major)"""
    if layout == FRAME_MAJOR:
        return flat.view(frames, lanes, width)
    return flat.view(lanes, frames, width).permute(1, 0, 2)


def shard_flat(flat: torch.Tensor, frames: int, lanes: int, lo: int, hi: int, layout: int,
               width: int = 1) -> torch.Tensor:
    """flat buffer of all lanes -> contiguous flat buffer of lanes [lo, hi) in the same layout"""
    v = _as_tlw(flat, frames, lanes, width, layout)[:, lo:hi]
    if layout == FRAME_MAJOR:
        return v.contiguous().view(-1)
    return v.permute(1, 0, 2).contiguous().view(-1)


def unshard_into(dst_flat: torch.Tensor, part: torch.Tensor, frames: int, lanes: int, lo: int, hi: int,
                 layout: int, width: int = 1) -> None:
    n = hi - lo
    if layout == FRAME_MAJOR:
        dst_flat.view(frames, lanes, width)[:, lo:hi] = part.view(frames, n, width)
    else:
        dst_flat.view(lanes, frames, width)[lo:hi] = part.view(n, frames, width)


def shard_state(words: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    """SoA state [W, lanes] -> [W, hi-lo] (contiguous copy)"""
    return words[:, lo:hi].contiguous()


def scatter_lanes(flat: Optional[torch.Tensor], frames: int, lanes: int, layout: int, width: int = 1,
                  src: int = 0, dtype=None, device=None, group=None) -> torch.Tensor:
    """Root holds `flat` for all lanes; every rank receives its lane block (same layout)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    blocks = all_blocks(world, lanes)
    lo, hi = blocks[rank]
    if rank == src:
        dtype, device = flat.dtype, flat.device
        parts = [shard_flat(flat, frames, lanes, a, b, layout, width) for a, b in blocks]
    else:
        parts = None
    out = torch.empty(frames * (hi - lo) * width, dtype=dtype, device=device)
    if world == 1:
        out.copy_(parts[0])
        return out
    # blocks may differ in size: one batch of point-to-point operations (NCCL runs them as a single
    # grouped launch over NVLink / NVSwitch)
    if rank == src:
        ops = [dist.P2POp(dist.isend, parts[r], r, group) for r in range(world) if r != src]
        out.copy_(parts[src])
    else:
        ops = [dist.P2POp(dist.irecv, out, src, group)]
    for q in dist.batch_isend_irecv(ops):
        q.wait()
    return out


def gather_lanes(part: torch.Tensor, frames: int, lanes: int, layout: int, width: int = 1, dst: int = 0,
                 group=None) -> Optional[torch.Tensor]:
    """Inverse of scatter_lanes: root receives the full flat buffer, others None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    blocks = all_blocks(world, lanes)
    if rank != dst:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, part, dst, group)]):
            q.wait()
        return None
    full = torch.empty(frames * lanes * width, dtype=part.dtype, device=part.device)
    bufs = [part if r == dst else torch.empty(frames * (b - a) * width, dtype=part.dtype, device=part.device)
            for r, (a, b) in enumerate(blocks)]
    ops = [dist.P2POp(dist.irecv, bufs[r], r, group) for r in range(world) if r != dst]
    for q in (dist.batch_isend_irecv(ops) if ops else []):
        q.wait()
    for r, (a, b) in enumerate(blocks):
        unshard_into(full, bufs[r], frames, lanes, a, b, layout, width)
    return full


class Comm:
    """The C ABI's communicator (``idsp_b200_comm_*``, raw NCCL over NVLink / NVSwitch on the ctx stream):
    what a non-Python caller uses to shard lanes.  Here the 128-byte id travels through
    ``torch.distributed`` (any out-of-band channel works); the data plane itself does not touch torch.

    ``scatter_lanes(full, frames, lanes, layout, width)`` -> this rank's lane block (root passes the full
    buffer, the others ``None``); ``gather_lanes(part, full, ...)`` is the inverse.  Lane-major blocks are
    sent / received in place, frame-major blocks are packed with one strided device copy per peer."""

    def __init__(self, device: int, group=None, own_stream: bool = False):
        """own_stream: give the communicator its own ctx / CUDA stream so that transfers overlap kernels running
        on the default ctx; order the two with ``compute_after()`` / ``after_compute()``"""
        self.device = int(device)
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:  # single process: a one-rank communicator (never touches NCCL)
            self.rank, self.world = 0, 1
        from .engine import Context

        self._ctx = Context(self.device, use_torch_stream=False) if own_stream else default_context(self.device)
        self._L = _lib.lib()
        ident = [None]
        if self.rank == 0 and self.world > 1:
            buf = (C.c_ubyte * 128)()
            _lib.check(self._L.idsp_b200_comm_unique_id(buf))
            ident = [bytes(buf)]
        if self.world > 1:
            dist.broadcast_object_list(ident, src=0, group=group)
        h = C.c_void_p()
        idbuf = (C.c_ubyte * 128).from_buffer_copy(ident[0]) if ident[0] is not None else None
        _lib.check(self._L.idsp_b200_comm_init(self._ctx._h, self.world, self.rank, idbuf, C.byref(h)))
        self._h = h

    def lane_block(self, lanes: int) -> Tuple[int, int]:
        lo, hi = C.c_size_t(), C.c_size_t()
        _lib.check(self._L.idsp_b200_lane_block(lanes, self.world, self.rank, 32, C.byref(lo), C.byref(hi)))
        return int(lo.value), int(hi.value)

    def scatter_lanes(self, full: Optional[torch.Tensor], frames: int, lanes: int, layout: int, width: int = 1,
                      root: int = 0, dtype=torch.int32, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        lo, hi = self.lane_block(lanes)
        if full is not None:
            dtype = full.dtype
        if out is None:
            out = torch.empty(frames * (hi - lo) * width, dtype=dtype, device=f"cuda:{self.device}")
        eb = out.element_size() * width
        _lib.check(self._L.idsp_scatter_lanes(self._h, None if full is None else C.c_void_p(full.data_ptr()),
                                              C.c_void_p(out.data_ptr()), frames, lanes, eb, layout, root))
        return out

    def gather_lanes(self, part: torch.Tensor, full: Optional[torch.Tensor], frames: int, lanes: int, layout: int,
                     width: int = 1, root: int = 0) -> Optional[torch.Tensor]:
        if self.rank == root and full is None:
            full = torch.empty(frames * lanes * width, dtype=part.dtype, device=part.device)
        eb = part.element_size() * width
        _lib.check(self._L.idsp_gather_lanes(self._h, C.c_void_p(part.data_ptr()),
                                             None if full is None else C.c_void_p(full.data_ptr()),
                                             frames, lanes, eb, layout, root))
        return full

    # ---- pipelined edges: point-to-point pieces on the communicator's stream
    def group(self):
        comm = self

        class _G:
            def __enter__(self_):
                _lib.check(comm._L.idsp_comm_group_begin(comm._h))

            def __exit__(self_, *a):
                _lib.check(comm._L.idsp_comm_group_end(comm._h))

        return _G()

    def send(self, t: torch.Tensor, peer: int):
        _lib.check(self._L.idsp_comm_send(self._h, C.c_void_p(t.data_ptr()), t.numel() * t.element_size(), peer))

    def recv(self, t: torch.Tensor, peer: int):
        _lib.check(self._L.idsp_comm_recv(self._h, C.c_void_p(t.data_ptr()), t.numel() * t.element_size(), peer))

    def compute_after(self, compute_ctx=None):
        """kernels queued on the compute ctx from now on start after everything queued on the communicator so far"""
        cc = compute_ctx or default_context(self.device)
        _lib.check(self._L.idsp_b200_stream_wait(cc._h, self._ctx._h))

    def after_compute(self, compute_ctx=None):
        """transfers queued from now on start after the kernels queued on the compute ctx so far"""
        cc = compute_ctx or default_context(self.device)
        _lib.check(self._L.idsp_b200_stream_wait(self._ctx._h, cc._h))

    def sync(self):
        self._ctx.sync()

    def broadcast(self, buf: torch.Tensor, root: int = 0) -> torch.Tensor:
        _lib.check(self._L.idsp_broadcast(self._h, C.c_void_p(buf.data_ptr()), buf.numel() * buf.element_size(), root))
        return buf

    def close(self):
        if getattr(self, "_h", None):
            self._L.idsp_b200_comm_free(self._h)
            self._h = None


class PeerBuffer:
    """A device buffer owned by rank `owner` and mapped into every process of the group.

    All ranks call the constructor (collective: the 64-byte handle of ``idsp_b200_ipc_export`` is
    broadcast).  ``view(offset, numel)`` gives the range as a kernel output (`PeerView`); on the
    owner ``tensor()`` is a torch view of the whole buffer.  ``close()`` is collective too.
    Lane-major results are the natural fit: the lane block [lo, hi) of a [lanes][frames * width]
    buffer is the contiguous range starting at lo * frames * width."""

    def __init__(self, numel: int, dtype: torch.dtype, device: int, owner: int = 0, group=None):
        self.numel, self.dtype, self.device, self.owner, self.group = int(numel), dtype, int(device), owner, group
        self.rank = dist.get_rank(group)
        self._ctx = default_context(self.device)
        self._L = _lib.lib()
        self.itemsize = torch.empty(0, dtype=dtype).element_size()
        p = C.c_void_p()
        handle = [None]
        if self.rank == owner:
            _lib.check(self._L.idsp_b200_malloc(self._ctx._h, self.numel * self.itemsize, C.byref(p)))
            buf = (C.c_ubyte * 64)()
            _lib.check(self._L.idsp_b200_ipc_export(self._ctx._h, p, buf))
            handle = [bytes(buf)]
        dist.broadcast_object_list(handle, src=owner, group=group)
        if self.rank != owner:
            buf = (C.c_ubyte * 64).from_buffer_copy(handle[0])
            _lib.check(self._L.idsp_b200_ipc_open(self._ctx._h, buf, C.byref(p)))
        self.ptr = p.value

    def view(self, offset: int, numel: int) -> PeerView:
        assert 0 <= offset and offset + numel <= self.numel
        return PeerView(self.ptr + offset * self.itemsize, numel, self.dtype)

    def tensor(self) -> torch.Tensor:
        """the owner's torch view of the buffer (no copy)"""
        assert self.rank == self.owner, "only the owning rank has a local view"

        class _Iface:
            pass

        o = _Iface()
        typestr = {torch.int32: "<i4", torch.float32: "<f4", torch.int64: "<i8", torch.float64: "<f8",
                   torch.int16: "<i2", torch.int8: "|i1"}[self.dtype]
        o.__cuda_array_interface__ = {"shape": (self.numel,), "typestr": typestr, "data": (self.ptr, False), "version": 2}
        self._keep = o
        return torch.as_tensor(o, device=f"cuda:{self.device}")

    def close(self) -> None:
        """collective: importers unmap first, then the owner frees"""
        if self.ptr is None:
            return
        torch.cuda.synchronize(self.device)  # kernels that still write into the mapping must have finished
        if self.rank != self.owner:
            _lib.check(self._L.idsp_b200_ipc_close(self._ctx._h, C.c_void_p(self.ptr)))
        dist.barrier(group=self.group)
        if self.rank == self.owner:
            _lib.check(self._L.idsp_b200_mfree(self._ctx._h, C.c_void_p(self.ptr)))
        self.ptr = None
