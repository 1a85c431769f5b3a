#!/usr/bin/env python3
"""Host <-> device copy rates under contention, by host-memory flavour (what bounds the `_host` entry points at
N > 1).  Every rank copies 256 MiB in both directions at once, all ranks at the same time (barrier before every
repetition); the slowest rank's best repetition is printed per flavour:

  pinned      cudaHostAlloc(default)                       (what torch.Tensor.pin_memory() gives)
  wc          cudaHostAlloc(cudaHostAllocWriteCombined)    for the host -> device side
  hugepage    mmap + madvise(MADV_HUGEPAGE) + cudaHostRegister

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py
"""
import ctypes
import mmap
import os

import torch
import torch.distributed as dist

MB = 256
N = MB << 20


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(N, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def host(flavour):
        p = ctypes.c_void_p()
        if flavour == "hugepage":
            m = mmap.mmap(-1, N + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
            addr = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
            libc = ctypes.CDLL("libc.so.6", use_errno=True)
            libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
            r = libc.madvise(addr, N, 14)  # MADV_HUGEPAGE
            ctypes.memset(addr, 1, N)  # touch
            e = rt.cudaHostRegister(addr, N, 1)  # portable
            assert e == 0, f"cudaHostRegister {e}"
            return addr, (m, r)
        e = rt.cudaHostAlloc(ctypes.byref(p), N, 4 if flavour == "wc" else 0)
        assert e == 0, f"cudaHostAlloc {e}"
        ctypes.memset(p.value, 1, N)
        return p.value, None

    for flavour in ("pinned", "wc", "hugepage"):
        try:
            h_in, keep1 = host(flavour)
            h_out, keep2 = host("pinned" if flavour == "wc" else flavour)  # WC memory is for the H2D side only
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"{flavour:9s} unavailable: {e}")
            continue
        best = [0.0, 0.0]
        for _ in range(5):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ev[0].record(s1)
            rt.cudaMemcpyAsync(d_in.data_ptr(), h_in, N, 1, ctypes.c_void_p(s1.cuda_stream))
            ev[1].record(s1)
            ev[2].record(s2)
            rt.cudaMemcpyAsync(h_out, d_out.data_ptr(), N, 2, ctypes.c_void_p(s2.cuda_stream))
            ev[3].record(s2)
            torch.cuda.synchronize()
            best[0] = max(best[0], N / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9)
            best[1] = max(best[1], N / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9)
        t = torch.tensor(best, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists("/sys/kernel/mm/transparent_hugepage/enabled") else "?"
            print(f"N={world} {flavour:9s} h2d {float(t[0]):6.1f} GB/s  d2h {float(t[1]):6.1f} GB/s per GPU, both directions at once (slowest rank)" + (f"  [THP: {thp}]" if flavour == "hugepage" else ""), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
