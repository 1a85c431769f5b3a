// lane_kernels.cuh -- generic "one filter lane per thread" streaming kernels.
//
// The recurrences (biquad, lowpass, lock-in) are serial in time per lane and
// independent across lanes (dsp-process/src/compose.rs:468-513), so a thread owns
// a lane, keeps its state in registers for the whole call and walks the time
// axis.  Two memory layouts (dsp-process/src/view.rs):
//   frame-major  flat[t*lanes + l]  -> a warp reads 32 adjacent lanes of one frame
//                                      (coalesced), U frames in flight per thread
//   lane-major   flat[l*frames + t] -> a warp transposes 32x32 tiles through shared
//                                      memory so global accesses stay coalesced
// These kernels work for every Op/type/shape; tma_kernels.cuh holds the
// TMA-pipelined specialisations used when alignment allows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "common.cuh"

namespace idsp {

// streaming loads/stores: coherent (x and y may alias), no L1 allocation
template <class T> __device__ __forceinline__ T ld_stream(const T *p) { return *p; }
template <> __device__ __forceinline__ int32_t ld_stream<int32_t>(const int32_t *p) {
    int32_t v;
    asm volatile("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <> __device__ __forceinline__ float ld_stream<float>(const float *p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <class T> __device__ __forceinline__ void st_stream(T *p, T v) { *p = v; }
template <> __device__ __forceinline__ void st_stream<int32_t>(int32_t *p, int32_t v) {
    asm volatile("st.global.L1::no_allocate.s32 [%0], %1;" ::"l"(p), "r"(v));
}
template <> __device__ __forceinline__ void st_stream<float>(float *p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v));
}
template <> __device__ __forceinline__ void st_stream<int2>(int2 *p, int2 v) {
    asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y));
}

// ---------------------------------------------------------------- frame-major
template <class Op, int U>
__global__ void __launch_bounds__(128)
lanes_fm_kernel(typename Op::Params p, const typename Op::In *x, typename Op::Out *y,
                size_t frames, size_t lanes, size_t sstride) {
    using In = typename Op::In;
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    Op op;
    op.load(p, lane, sstride);
    const In *xp = x + lane;
    typename Op::Out *yp = y + lane;
    const size_t nblk = frames / U;
    In cur[U], nxt[U];
    if (nblk) {
#pragma unroll
        for (int u = 0; u < U; u++) cur[u] = ld_stream(xp + (size_t)u * lanes);
    }
    for (size_t b = 0; b < nblk; b++) {
        const In *xn = xp + (b + 1) * U * lanes;
        if (b + 1 < nblk) {
#pragma unroll
            for (int u = 0; u < U; u++) nxt[u] = ld_stream(xn + (size_t)u * lanes);
        }
        typename Op::Out *yb = yp + b * U * lanes;
#pragma unroll
        for (int u = 0; u < U; u++) st_stream(yb + (size_t)u * lanes, op.step(p, cur[u]));
#pragma unroll
        for (int u = 0; u < U; u++) cur[u] = nxt[u];
    }
    for (size_t t = nblk * U; t < frames; t++)
        st_stream(yp + t * lanes, op.step(p, ld_stream(xp + t * lanes)));
    op.store(p, lane, sstride);
}

// ---------------------------------------------------------------- lane-major
// One warp = 32 lanes; 32x32 tiles transposed through padded shared memory.
template <class Op, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
lanes_lm_kernel(typename Op::Params p, const typename Op::In *x, typename Op::Out *y,
                size_t frames, size_t lanes, size_t sstride) {
    using In = typename Op::In;
    using Out = typename Op::Out;
    __shared__ In tin[WARPS][32][33];
    __shared__ Out tout[WARPS][32][33];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const size_t lane0 = ((size_t)blockIdx.x * WARPS + w) * 32;
    if (lane0 >= lanes) return;
    const int nl = (int)((lanes - lane0) < 32 ? (lanes - lane0) : 32);  // lanes in this warp
    const size_t lane = lane0 + l;
    const bool active = l < nl;
    Op op;
    if (active) op.load(p, lane, sstride);
    In r[32];
    auto fetch = [&](size_t t0) {
        const bool tok = t0 + l < frames;
#pragma unroll
        for (int i = 0; i < 32; i++)
            if (i < nl && tok) r[i] = ld_stream(x + (lane0 + i) * frames + t0 + l);
    };
    if (frames) fetch(0);
    for (size_t t0 = 0; t0 < frames; t0 += 32) {
        const int nt = (int)((frames - t0) < 32 ? (frames - t0) : 32);
#pragma unroll
        for (int i = 0; i < 32; i++) tin[w][i][l] = r[i];
        __syncwarp();
        if (t0 + 32 < frames) fetch(t0 + 32);
        if (active) {
            for (int j = 0; j < nt; j++) tout[w][l][j] = op.step(p, tin[w][l][j]);
        }
        __syncwarp();
        if (l < nt) {
#pragma unroll
            for (int i = 0; i < 32; i++)
                if (i < nl) st_stream(y + (lane0 + i) * frames + t0 + l, tout[w][i][l]);
        }
        __syncwarp();
    }
    if (active) op.store(p, lane, sstride);
}

// ---------------------------------------------------------------- lane-major, wide tiles
// For 1-, 2- and 8-byte samples (the 4-byte ones have the TMA kernels): a warp moves tiles of 128
// bytes per lane with 16-byte accesses (thread t handles pieces t, t+32, ...: 8 consecutive
// pieces = 128 contiguous bytes of one lane row), transposes them through shared memory with the
// piece index XOR-swizzled by (lane & 7) -- conflict free for the cooperative side and for the
// per-lane side -- and prefetches the next tile into registers.  Results overwrite the tile in
// place and leave the same way.  Needs 16-byte aligned lane rows (frames * sizeof(T) % 16 == 0).
template <class T> struct Piece16 {
    static constexpr int N = 16 / (int)sizeof(T);
    __device__ __forceinline__ static void unpack(const int4 &v, T *e) {
        const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
        if constexpr (sizeof(T) == 8) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const uint64_t u = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
                if constexpr (std::is_same<T, double>::value) e[i] = __longlong_as_double((long long)u);
                else e[i] = (T)u;
            }
        } else {
            constexpr int PW = 4 / (int)sizeof(T);  // elements per 32-bit word
#pragma unroll
            for (int i = 0; i < N; i++) e[i] = (T)(w[i / PW] >> (8 * sizeof(T) * (i % PW)));
        }
    }
    __device__ __forceinline__ static int4 pack(const T *e) {
        uint32_t w[4] = {0, 0, 0, 0};
        if constexpr (sizeof(T) == 8) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                uint64_t u;
                if constexpr (std::is_same<T, double>::value) u = (uint64_t)__double_as_longlong(e[i]);
                else u = (uint64_t)e[i];
                w[2 * i] = (uint32_t)u;
                w[2 * i + 1] = (uint32_t)(u >> 32);
            }
        } else {
            constexpr int PW = 4 / (int)sizeof(T);
            using UT = typename std::make_unsigned<T>::type;
#pragma unroll
            for (int i = 0; i < N; i++) w[i / PW] |= (uint32_t)(UT)e[i] << (8 * sizeof(T) * (i % PW));
        }
        return make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
    }
};

template <class Op, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
lanes_lm_wide_kernel(typename Op::Params p, const typename Op::In *x, typename Op::Out *y, size_t frames,
                     size_t lanes, size_t sstride) {
    using T = typename Op::In;
    static_assert(sizeof(typename Op::In) == sizeof(typename Op::Out), "in-place tile needs equal sample sizes");
    constexpr int EPP = 16 / (int)sizeof(T);  // samples per piece
    constexpr int EPT = 8 * EPP;              // frames per tile (128 bytes per lane)
    __shared__ int4 tile[WARPS][32 * 8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const size_t lane0 = ((size_t)blockIdx.x * WARPS + w) * 32;
    if (lane0 >= lanes) return;
    const int nl = (int)((lanes - lane0) < 32 ? (lanes - lane0) : 32);
    const bool active = l < nl;
    Op op;
    if (active) op.load(p, lane0 + l, sstride);
    const size_t ntiles = frames / EPT;
    int4 pre[8];
    auto fetch = [&](size_t ti) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int c = l + 32 * k, lane = c >> 3, q = c & 7;
            pre[k] = lane < nl ? *reinterpret_cast<const int4 *>(x + (lane0 + lane) * frames + ti * EPT + q * EPP)
                               : make_int4(0, 0, 0, 0);
        }
    };
    if (ntiles) fetch(0);
    for (size_t ti = 0; ti < ntiles; ti++) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int c = l + 32 * k, lane = c >> 3, q = c & 7;
            tile[w][lane * 8 + (q ^ (lane & 7))] = pre[k];
        }
        __syncwarp();
        if (ti + 1 < ntiles) fetch(ti + 1);
        if (active) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                int4 &slot = tile[w][l * 8 + (q ^ (l & 7))];
                T e[EPP];
                typename Op::Out o[EPP];
                Piece16<T>::unpack(slot, e);
#pragma unroll
                for (int i = 0; i < EPP; i++) o[i] = op.step(p, e[i]);
                slot = Piece16<typename Op::Out>::pack(o);
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int c = l + 32 * k, lane = c >> 3, q = c & 7;
            if (lane < nl)
                *reinterpret_cast<int4 *>(y + (lane0 + lane) * frames + ti * EPT + q * EPP) =
                    tile[w][lane * 8 + (q ^ (lane & 7))];
        }
    }
    if (active) {
        for (size_t t = ntiles * EPT; t < frames; t++)
            y[(lane0 + l) * frames + t] = op.step(p, x[(lane0 + l) * frames + t]);
        op.store(p, lane0 + l, sstride);
    }
}

template <class Op>
static int launch_lanes(idsp_ctx *ctx, const typename Op::Params &p, const typename Op::In *x,
                        typename Op::Out *y, size_t frames, size_t lanes, size_t sstride,
                        int layout) {
    if (lanes == 0) return IDSP_OK;
    if (layout == IDSP_FRAME_MAJOR) {
        constexpr int U = sizeof(typename Op::In) >= 8 ? 8 : 16;
        unsigned grid = (unsigned)((lanes + 63) / 64);
        lanes_fm_kernel<Op, U><<<grid, 64, 0, ctx->stream>>>(p, x, y, frames, lanes, sstride);
        IDSP_KERNEL_FAMILY(ctx, "generic frame-major");
    } else if constexpr (sizeof(typename Op::In) == sizeof(typename Op::Out) && sizeof(typename Op::In) != 4 &&
                         (std::is_integral<typename Op::In>::value || std::is_same<typename Op::In, double>::value)) {
        constexpr int WARPS = 4;
        const bool wide = ctx->policy != 1 && (frames * sizeof(typename Op::In)) % 16 == 0 &&
                          ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0 && frames * sizeof(typename Op::In) >= 128;
        if (wide) {
            unsigned grid = (unsigned)((lanes + WARPS * 32 - 1) / (WARPS * 32));
            lanes_lm_wide_kernel<Op, WARPS><<<grid, WARPS * 32, 0, ctx->stream>>>(p, x, y, frames, lanes, sstride);
            IDSP_KERNEL_FAMILY(ctx, "generic lane-major wide tiles");
        } else {
            unsigned grid = (unsigned)((lanes + 2 * 32 - 1) / (2 * 32));
            lanes_lm_kernel<Op, 2><<<grid, 2 * 32, 0, ctx->stream>>>(p, x, y, frames, lanes, sstride);
            IDSP_KERNEL_FAMILY(ctx, "generic lane-major");
        }
    } else {
        constexpr int WARPS = 2;
        unsigned grid = (unsigned)((lanes + WARPS * 32 - 1) / (WARPS * 32));
        lanes_lm_kernel<Op, WARPS><<<grid, WARPS * 32, 0, ctx->stream>>>(p, x, y, frames, lanes,
                                                                        sstride);
        IDSP_KERNEL_FAMILY(ctx, "generic lane-major");
    }
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

}  // namespace idsp

