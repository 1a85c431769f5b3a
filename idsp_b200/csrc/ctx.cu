// ctx.cu -- context lifetime, error reporting, host-buffer streaming.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void idsp_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *idsp_b200_last_error(void) { return g_err; }
extern "C" int idsp_b200_version(void) { return IDSP_B200_VERSION; }

static int ctx_create(int device, cudaStream_t stream, bool own, idsp_ctx **out) {
    if (!out) {
        idsp_set_error("idsp_b200_init: out is null");
        return IDSP_EINVAL;
    }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        idsp_set_error("idsp_b200_init: no CUDA device (%s)", cudaGetErrorString(e));
        return IDSP_ENODEV;
    }
    if (device < 0 || device >= n) {
        idsp_set_error("idsp_b200_init: device %d out of range (0..%d)", device, n - 1);
        return IDSP_EINVAL;
    }
    cudaDeviceProp prop;
    IDSP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        // sm_100a-only binary: there is deliberately no other code path
        idsp_set_error("idsp_b200_init: device %d is sm_%d%d, this library is built for sm_100a only",
                       device, prop.major, prop.minor);
        return IDSP_ENODEV;
    }
    IDSP_CUDA(cudaSetDevice(device));
    idsp_ctx *c = new idsp_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->own_stream = own;
    if (own) {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            idsp_set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
            delete c;
            return IDSP_ECUDA;
        }
    } else {
        c->stream = stream;
    }
    *out = c;
    return IDSP_OK;
}

extern "C" int idsp_b200_init(int device, idsp_ctx **out) {
    return ctx_create(device, nullptr, true, out);
}
extern "C" int idsp_b200_init_on_stream(int device, void *cuda_stream, idsp_ctx **out) {
    return ctx_create(device, (cudaStream_t)cuda_stream, false, out);
}

static void free_staging(idsp_ctx *c) {
    for (int i = 0; i < IDSP_HOST_RING; i++) {
        if (c->dev_in[i]) cudaFree(c->dev_in[i]);
        if (c->dev_out[i]) cudaFree(c->dev_out[i]);
        c->dev_in[i] = c->dev_out[i] = nullptr;
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
        c->ev_h2d[i] = c->ev_k[i] = c->ev_d2h[i] = nullptr;
    }
    if (c->dev_state) cudaFree(c->dev_state);
    c->dev_state = nullptr;
    if (c->dev_scratch) cudaFree(c->dev_scratch);
    c->dev_scratch = nullptr;
    if (c->ev_order) cudaEventDestroy(c->ev_order);
    c->ev_order = nullptr;
    c->dev_in_bytes = c->dev_out_bytes = c->dev_state_bytes = c->dev_scratch_bytes = 0;
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    c->s_h2d = c->s_d2h = nullptr;
}

int idsp_scratch(idsp_ctx *ctx, size_t bytes, void **ptr) {
    if (ctx->dev_scratch_bytes < bytes) {
        if (ctx->dev_scratch) {
            IDSP_CUDA(cudaStreamSynchronize(ctx->stream));  // queued kernels may still use the old buffer
            cudaFree(ctx->dev_scratch);
            ctx->dev_scratch = nullptr;
            ctx->dev_scratch_bytes = 0;
        }
        cudaError_t e = cudaMalloc(&ctx->dev_scratch, bytes);
        if (e != cudaSuccess) {
            idsp_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
            return IDSP_ENOMEM;
        }
        ctx->dev_scratch_bytes = bytes;
    }
    *ptr = ctx->dev_scratch;
    return IDSP_OK;
}

extern "C" void idsp_b200_free(idsp_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_staging(ctx);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int idsp_b200_sync(idsp_ctx *ctx) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    IDSP_CUDA(cudaStreamSynchronize(ctx->stream));
    return IDSP_OK;
}

extern "C" int idsp_b200_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) {
        idsp_set_error("idsp_b200_host_alloc: ptr is null");
        return IDSP_EINVAL;
    }
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        idsp_set_error("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
        return IDSP_ENOMEM;
    }
    return IDSP_OK;
}
extern "C" void idsp_b200_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

extern "C" uint64_t idsp_b200_launch_count(const idsp_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" const char *idsp_b200_last_kernel(const idsp_ctx *ctx) {
    return ctx && ctx->last_kernel ? ctx->last_kernel : "";
}

// ---------------------------------------------------------------------------
// Peer memory (CUDA IPC): a plain cudaMalloc allocation (exportable, unlike a sub-block of a caching
// allocator), its handle, and the importing side.  Opening enables peer access lazily.
// ---------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == IDSP_IPC_HANDLE_BYTES, "handle size");
extern "C" int idsp_b200_malloc(idsp_ctx *ctx, size_t bytes, void **ptr) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    IDSP_CHECK_ARG(ptr != nullptr, "ptr is null");
    *ptr = nullptr;
    cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        idsp_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        return IDSP_ENOMEM;
    }
    return IDSP_OK;
}
extern "C" int idsp_b200_mfree(idsp_ctx *ctx, void *ptr) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    if (ptr) IDSP_CUDA(cudaFree(ptr));
    return IDSP_OK;
}
extern "C" int idsp_b200_ipc_export(idsp_ctx *ctx, const void *ptr, unsigned char handle[IDSP_IPC_HANDLE_BYTES]) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    IDSP_CHECK_ARG(ptr != nullptr && handle != nullptr, "ptr/handle is null");
    cudaIpcMemHandle_t h;
    IDSP_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(ptr)));
    memcpy(handle, &h, sizeof(h));
    return IDSP_OK;
}
extern "C" int idsp_b200_ipc_open(idsp_ctx *ctx, const unsigned char handle[IDSP_IPC_HANDLE_BYTES], void **ptr) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    IDSP_CHECK_ARG(ptr != nullptr && handle != nullptr, "ptr/handle is null");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    *ptr = nullptr;
    IDSP_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return IDSP_OK;
}
extern "C" int idsp_b200_ipc_close(idsp_ctx *ctx, void *ptr) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    if (ptr) IDSP_CUDA(cudaIpcCloseMemHandle(ptr));
    return IDSP_OK;
}

extern "C" int idsp_b200_memcpy(idsp_ctx *ctx, void *dst, const void *src, size_t bytes, int kind) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    if (bytes == 0) return IDSP_OK;
    IDSP_CHECK_ARG(dst != nullptr && src != nullptr, "dst/src is null");
    IDSP_CHECK_ARG(kind >= 0 && kind <= 2, "kind must be 0 (h2d), 1 (d2h) or 2 (d2d)");
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    IDSP_CUDA(cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
    return IDSP_OK;
}
extern "C" int idsp_b200_memset(idsp_ctx *ctx, void *ptr, int value, size_t bytes) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    if (bytes == 0) return IDSP_OK;
    IDSP_CHECK_ARG(ptr != nullptr, "ptr is null");
    IDSP_CUDA(cudaMemsetAsync(ptr, value, bytes, ctx->stream));
    return IDSP_OK;
}

extern "C" int idsp_b200_stream_wait(idsp_ctx *waiter, idsp_ctx *signal) {
    int r = idsp_use_device(signal);
    if (r) return r;
    IDSP_CHECK_ARG(waiter != nullptr, "waiter is null");
    IDSP_CHECK_ARG(waiter->device == signal->device, "both contexts must be on the same device");
    if (!signal->ev_order) IDSP_CUDA(cudaEventCreateWithFlags(&signal->ev_order, cudaEventDisableTiming));
    IDSP_CUDA(cudaEventRecord(signal->ev_order, signal->stream));
    IDSP_CUDA(cudaStreamWaitEvent(waiter->stream, signal->ev_order, 0));
    return IDSP_OK;
}

extern "C" int idsp_b200_set_kernel_policy(idsp_ctx *ctx, int policy) {
    if (!ctx || policy < 0 || policy > 3) {
        idsp_set_error("idsp_b200_set_kernel_policy: bad argument");
        return IDSP_EINVAL;
    }
    ctx->policy = policy;
    return IDSP_OK;
}

// ---------------------------------------------------------------------------
// Host streaming: x/y/state live in host memory.  The frame (frame-major) or
// lane (lane-major) axis is cut into chunks of ~32 MiB that rotate through a ring of
// IDSP_HOST_RING device buffers: the H2D copy of chunk i+1.. overlaps the kernel on
// chunk i and the D2H copy of chunk i-1 (three streams, events), so both PCIe
// directions stay busy.  If the caller's buffers are pinned (cudaHostAlloc /
// cudaHostRegister / torch pin_memory) the copy engines DMA them directly; pageable
// buffers still work (the driver stages them).
// ---------------------------------------------------------------------------
static int ensure_dev(void **p, size_t *have, size_t need) {
    if (*have >= need) return IDSP_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) {
        idsp_set_error("cudaMalloc(%zu): %s", need, cudaGetErrorString(e));
        return IDSP_ENOMEM;
    }
    *have = need;
    return IDSP_OK;
}

static int ensure_ring(void **bufs, size_t *have, size_t need) {
    if (*have >= need) return IDSP_OK;
    for (int i = 0; i < IDSP_HOST_RING; i++) {
        if (bufs[i]) cudaFree(bufs[i]);
        bufs[i] = nullptr;
    }
    *have = 0;
    for (int i = 0; i < IDSP_HOST_RING; i++) {
        cudaError_t e = cudaMalloc(&bufs[i], need ? need : 256);
        if (e != cudaSuccess) {
            idsp_set_error("cudaMalloc(%zu): %s", need, cudaGetErrorString(e));
            return IDSP_ENOMEM;
        }
    }
    *have = need;
    return IDSP_OK;
}

int idsp_host_stream(idsp_ctx *ctx, const HostStreamSpec &spec, const void *x, void *y,
                     const HostStreamLaunch &launch) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    constexpr int NB = IDSP_HOST_RING;
    if (!ctx->s_h2d) {
        IDSP_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        IDSP_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < NB; i++) {
            IDSP_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
            IDSP_CUDA(cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
            IDSP_CUDA(cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming));
        }
    }
    // state blobs -> device (one allocation, 256 B aligned sub-blobs)
    size_t off[3] = {0, 0, 0}, tot = 0;
    for (int i = 0; i < spec.nblobs; i++) {
        off[i] = tot;
        tot += (spec.blobs[i].bytes + 255) & ~(size_t)255;
    }
    r = ensure_dev(&ctx->dev_state, &ctx->dev_state_bytes, tot ? tot : 256);
    if (r) return r;
    void *dblobs[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < spec.nblobs; i++) {
        dblobs[i] = (char *)ctx->dev_state + off[i];
        IDSP_CUDA(cudaMemcpyAsync(dblobs[i], spec.blobs[i].host, spec.blobs[i].bytes,
                                  cudaMemcpyHostToDevice, ctx->stream));
    }
    const bool fm = spec.layout == IDSP_FRAME_MAJOR;
    // chunk axis: frames (frame-major) or lanes (lane-major); unit = bytes per index
    const size_t n_axis = fm ? spec.frames : spec.lanes;
    const size_t other = fm ? spec.lanes : spec.frames;
    const size_t in_unit = other * spec.in_bytes_per_frame_lane;
    const size_t out_unit = other * spec.out_bytes_per_frame_lane;
    const size_t bigger = in_unit > out_unit ? in_unit : out_unit;
    // chunk size per direction: small enough that the un-overlapped first H2D / last D2H are a
    // few per cent of a call, large enough to keep PCIe efficient (IDSP_HOST_CHUNK_MB overrides)
    const char *cm = getenv("IDSP_HOST_CHUNK_MB");  // read per call: no shared mutable state between ctxs
    const size_t chunk_mb = cm && atoi(cm) > 0 ? (size_t)atoi(cm) : 32;
    const size_t target = chunk_mb << 20;
    size_t chunk = bigger ? target / bigger : n_axis;
    if (fm) {
        if (chunk > 512) chunk &= ~(size_t)511;  // whole tiles of the tiled kernels
    } else {
        chunk &= ~(size_t)127;  // keep lane chunks warp/CTA aligned
        if (chunk < 128) chunk = 128;
    }
    if (chunk < 1) chunk = 1;
    if (chunk > n_axis) chunk = n_axis;
    if (n_axis == 0) chunk = 0;
    r = ensure_ring(ctx->dev_in, &ctx->dev_in_bytes, chunk * in_unit);
    if (r) return r;
    r = ensure_ring(ctx->dev_out, &ctx->dev_out_bytes, chunk * out_unit);
    if (r) return r;
    const size_t nchunks = chunk ? (n_axis + chunk - 1) / chunk : 0;
    for (size_t c = 0; c < nchunks; c++) {
        const int b = (int)(c % NB);
        const size_t a0 = c * chunk;
        const size_t an = (n_axis - a0) < chunk ? (n_axis - a0) : chunk;
        // buffer b's previous kernel (chunk c-NB) must be done before its input is overwritten,
        // and its previous D2H must be done before the kernel overwrites its output
        if (c >= (size_t)NB) {
            IDSP_CUDA(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_k[b], 0));
            IDSP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[b], 0));
        }
        IDSP_CUDA(cudaMemcpyAsync(ctx->dev_in[b], (const char *)x + a0 * in_unit, an * in_unit,
                                  cudaMemcpyHostToDevice, ctx->s_h2d));
        IDSP_CUDA(cudaEventRecord(ctx->ev_h2d[b], ctx->s_h2d));
        IDSP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[b], 0));
        r = launch(dblobs, ctx->dev_in[b], ctx->dev_out[b], a0, an);
        if (r) {
            // copies into the caller's buffers may still be in flight: drain all three streams before
            // handing the error back (the caller may free x / y right away)
            cudaStreamSynchronize(ctx->s_h2d);
            cudaStreamSynchronize(ctx->stream);
            cudaStreamSynchronize(ctx->s_d2h);
            return r;
        }
        IDSP_CUDA(cudaEventRecord(ctx->ev_k[b], ctx->stream));
        IDSP_CUDA(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_k[b], 0));
        IDSP_CUDA(cudaMemcpyAsync((char *)y + a0 * out_unit, ctx->dev_out[b], an * out_unit,
                                  cudaMemcpyDeviceToHost, ctx->s_d2h));
        IDSP_CUDA(cudaEventRecord(ctx->ev_d2h[b], ctx->s_d2h));
    }
    for (int i = 0; i < spec.nblobs; i++) {
        if (spec.blobs[i].writeback)
            IDSP_CUDA(cudaMemcpyAsync(spec.blobs[i].host, dblobs[i], spec.blobs[i].bytes,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    }
    IDSP_CUDA(cudaStreamSynchronize(ctx->s_d2h));
    IDSP_CUDA(cudaStreamSynchronize(ctx->stream));
    IDSP_CUDA(cudaStreamSynchronize(ctx->s_h2d));
    return IDSP_OK;
}
