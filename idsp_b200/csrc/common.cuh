// common.cuh -- context, error plumbing and launch helpers shared by all .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/idsp_b200.h"

#define IDSP_HOST_RING 4

struct idsp_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    uint64_t launches;
    int policy;  // 0 auto, 1 generic, 2 TMA / tiled, 3 auto with the packed f32x2 HBF variant
    int sm_count;
    const char *last_kernel;  // family of the most recent launch (idsp_b200_last_kernel)
    // host streaming (the *_host entry points): ring of device chunk buffers
    cudaStream_t s_h2d, s_d2h;
    void *dev_in[IDSP_HOST_RING], *dev_out[IDSP_HOST_RING];
    size_t dev_in_bytes, dev_out_bytes;
    void *dev_state;
    size_t dev_state_bytes;
    // scratch for entry points that compose several kernels (stream-ordered reuse)
    void *dev_scratch;
    size_t dev_scratch_bytes;
    cudaEvent_t ev_h2d[IDSP_HOST_RING], ev_k[IDSP_HOST_RING], ev_d2h[IDSP_HOST_RING];
    cudaEvent_t ev_order;  // idsp_b200_stream_wait
};

void idsp_set_error(const char *fmt, ...);

#define IDSP_CHECK_ARG(cond, msg)                          \
    do {                                                   \
        if (!(cond)) {                                     \
            idsp_set_error("%s: %s", __func__, msg);       \
            return IDSP_EINVAL;                            \
        }                                                  \
    } while (0)

#define IDSP_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            idsp_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
            return IDSP_ECUDA;                                                       \
        }                                                                            \
    } while (0)

// after a kernel launch
#define IDSP_KERNEL_FAMILY(ctx, name) ((ctx)->last_kernel = (name))
#define IDSP_LAUNCHED(ctx)                                                           \
    do {                                                                             \
        (ctx)->launches++;                                                           \
        cudaError_t e_ = cudaGetLastError();                                         \
        if (e_ != cudaSuccess) {                                                     \
            idsp_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e_)); \
            return IDSP_ECUDA;                                                       \
        }                                                                            \
    } while (0)

// grows ctx->dev_scratch to at least `bytes` (ctx.cu); the previous buffer is released only after
// the work queued on the ctx stream has finished
int idsp_scratch(idsp_ctx *ctx, size_t bytes, void **ptr);

static inline int idsp_use_device(idsp_ctx *ctx) {
    if (!ctx) {
        idsp_set_error("null ctx");
        return IDSP_EINVAL;
    }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) {
        idsp_set_error("cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
        return IDSP_ECUDA;
    }
    return IDSP_OK;
}

// f32 DF1 biquad on a lane block of a larger SoA state (biquad.cu; used by the chain composition in hbf.cu)
int idsp_df1_f32_strided(idsp_ctx *ctx, const float *ba, float *state, const float *x, float *y, size_t frames,
                         size_t lanes, size_t sstride, int layout);

// Host streaming helper (ctx.cu): cuts the frame axis (frame-major) or the lane axis
// (lane-major) into chunks [a0, a0+an) and runs `launch(dev_blobs, dev_x, dev_y, a0, an)`
// per chunk with H2D / compute / D2H overlapped.  dev_x/dev_y hold only the chunk.
#include <functional>
struct HostStreamSpec {
    size_t frames;            // total frames
    size_t lanes;
    size_t in_bytes_per_frame_lane;   // bytes per lane per frame of x
    size_t out_bytes_per_frame_lane;  // bytes per lane per frame of y
    int layout;
    // state blobs (host <-> device), copied before/after
    struct Blob { void *host; size_t bytes; bool writeback; };
    Blob blobs[3];
    int nblobs;
};
typedef std::function<int(void **dev_blobs, const void *dx, void *dy, size_t a0, size_t an)> HostStreamLaunch;
int idsp_host_stream(idsp_ctx *ctx, const HostStreamSpec &spec, const void *x, void *y,
                     const HostStreamLaunch &launch);
