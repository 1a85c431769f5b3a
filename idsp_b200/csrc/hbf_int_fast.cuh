// hbf_int_fast.cuh -- shared-memory tiled HBF x2^K interpolation cascade (f32, lane-major).
//
// Mirror image of hbf_fast_scalar.cuh.  A CTA owns NL = 8 lanes and walks time in tiles of
// TI = 512 >> K input samples per lane (512 output samples per lane per tile):
//
//   x tile (128 B per lane) --register prefetch--> X rows [hist | TI]
//   stage 0 (TAPS[0], lowest rate) : X -> U_1           (shared memory -> shared memory)
//   stage s                        : U_s -> U_{s+1}
//   stage K-1 (highest rate)       : U_{K-1} -> output staging ring (2 x 2 KB per lane)
//   staging --TMA 1-D bulk store per lane row (2 KB, bulk_group)--> HBM
//
// Work item = (lane, R consecutive inputs of one stage) -> 2R outputs:
//   out[2n]   = ((x[n]+x[n-2M+1])*c0) + ((x[n-1]+x[n-2M+2])*c1) + ...   (src/hbf.rs:46-68, :220-224)
//   out[2n+1] = x[n-M+1]                                              (centre tap = identity)
// each op individually rounded, summed in the reference's order -> bit-exact.  Items are
// spread lane-fastest over the threads; rows have a pitch of 4*odd floats so every
// LDS.128 / STS.128 over 8 lanes is conflict free.  Rows are [history | tile]; tails move to
// heads after each tile (copy_within, src/hbf.rs:231); ABI state <-> those heads at entry/exit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "hbf_stages.cuh"
#include "tma_kernels.cuh"

namespace idsp {
namespace hfi {

constexpr int NL = 8;     // lanes per CTA
constexpr int NT = 128;   // threads per CTA
constexpr int TOUT = 512; // output samples per lane per tile

__host__ __device__ constexpr int up4(int v) { return (v + 3) & ~3; }
__host__ __device__ constexpr int oddpitch(int v) { return (up4(v) / 4) % 2 ? up4(v) : up4(v) + 4; }
// stage s of a x2^K cascade uses TAPS[s] (lowest rate first, src/hbf.rs:503-512)
__host__ __device__ constexpr int st_m(int s) { return hbf_m(s); }
__host__ __device__ constexpr int ti(int K) { return TOUT >> K; }                 // inputs per tile
__host__ __device__ constexpr int st_nin(int K, int s) { return ti(K) << s; }      // inputs of stage s per tile
__host__ __device__ constexpr int st_r(int K, int s) {                             // inputs per item
    return st_nin(K, s) / 16 >= 8 ? 8 : 4;
}
__host__ __device__ constexpr int hist(int s) { return up4(2 * st_m(s) - 1); }
__host__ __device__ constexpr int pitch(int K, int s) { return oddpitch(hist(s) + st_nin(K, s)); }
__host__ __device__ constexpr int off_u(int K, int s) {  // float offset of stage s's input rows
    int o = 0;
    for (int i = 0; i < s; i++) o += NL * pitch(K, i);
    return o;
}
constexpr int OUT_PITCH = oddpitch(TOUT);
__host__ __device__ constexpr int off_out(int K) { return off_u(K, K); }
__host__ __device__ constexpr int smem_floats(int K) { return off_out(K) + 2 * NL * OUT_PITCH; }
__host__ __device__ constexpr size_t smem_bytes(int K) { return (size_t)smem_floats(K) * 4; }
__host__ __device__ constexpr int st_word(int s) {  // ABI state word offset of stage s
    int w = 0;
    for (int i = 0; i < s; i++) w += 2 * st_m(i) - 1;
    return w;
}

__device__ __forceinline__ float4 lds128v(const float *p) {  // volatile: never narrowed by ptxas
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void bulk_store_1d(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}

// One item: inputs n0 .. n0+R-1 of row `row` = [H hist | n new]; writes 2R outputs to dst
template <int TI_, int R> struct IntItem {
    static constexpr int M = HbfTaps<TI_>::M;
    static constexpr int LEN = 2 * M - 1;
    static constexpr int H = up4(LEN);
    static constexpr int RO = H - LEN;
    static constexpr int W = up4(RO + R + LEN);
    __device__ __forceinline__ static void run(const float *row, int n0, float *dst) {
        float w[W];
#pragma unroll
        for (int j = 0; j < W / 4; j++) {
            float4 v = lds128v(row + n0 + 4 * j);
            w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
        }
        float o[2 * R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            // window of input n0+q: w[RO+q .. RO+q+2M-1]
            float acc = (w[RO + q + 2 * M - 1] + w[RO + q]) * HbfTaps<TI_>::c(0);
#pragma unroll
            for (int i = 1; i < M; i++)
                acc = acc + (w[RO + q + 2 * M - 1 - i] + w[RO + q + i]) * HbfTaps<TI_>::c(i);
            o[2 * q] = acc;
            o[2 * q + 1] = w[RO + q + M];
        }
#pragma unroll
        for (int j = 0; j < 2 * R / 4; j++)
            reinterpret_cast<float4 *>(dst)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
};

// Move the tails of the [hist | n new] input rows of stage s (all NL lanes) to their heads
// (copy_within, src/hbf.rs:231).  One 16-byte piece per thread, a whole row inside one warp:
// everything is read before anything is written (head and tail overlap when hist > n).
template <int K, int s>
__device__ __forceinline__ void carry_rows(float *sm, int warp, int lid) {
    constexpr int C = hist(s) / 4;
    constexpr int LP = 32 / C;  // lanes per warp pass
    static_assert(C <= 32, "row history too long for one warp");
    const int sub = lid / C, j = lid % C;
    int lane = warp * LP + sub;
    float *row = sm + off_u(K, s) + lane * pitch(K, s) + 4 * j;
    for (; lane - sub < NL; lane += (NT / 32) * LP, row += (NT / 32) * LP * pitch(K, s)) {  // warp-uniform trip count
        const bool act = sub < LP && lane < NL;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) v = lds128v(row + st_nin(K, s));
        __syncwarp();
        if (act) *reinterpret_cast<float4 *>(row) = v;
        __syncwarp();
    }
}

template <int K, int s> struct StageRun {
    __device__ __forceinline__ static void run(float *sm, int tid, int obuf) {
        constexpr int R = st_r(K, s);
        constexpr int ITEMS = NL * st_nin(K, s) / R;
        const float *U = sm + off_u(K, s);
        for (int idx = tid; idx < ITEMS; idx += NT) {
            const int lane = idx % NL, n0 = (idx / NL) * R;
            float *dst;
            if constexpr (s == K - 1) dst = sm + off_out(K) + (obuf * NL + lane) * OUT_PITCH + 2 * n0;
            else dst = sm + off_u(K, s + 1) + lane * pitch(K, s + 1) + hist(s + 1) + 2 * n0;
            IntItem<s, R>::run(U + lane * pitch(K, s), n0, dst);
        }
        // the input rows of the previous stage were consumed one barrier ago
        if constexpr (s >= 1) carry_rows<K, s - 1>(sm, tid >> 5, tid & 31);
    }
};

template <int K, int s, bool LOAD> struct StateIO {
    __device__ __forceinline__ static void run(float *sm, float *st, size_t sstride, size_t lane0, int nl, int tid) {
        if constexpr (s < K) {
            constexpr int LEN = 2 * st_m(s) - 1;
            float *stw = st + (size_t)st_word(s) * sstride + lane0;
            for (int idx = tid; idx < LEN * NL; idx += NT) {
                const int lane = idx % NL, w = idx / NL;
                if (lane >= nl) continue;
                float *p = sm + off_u(K, s) + lane * pitch(K, s) + (hist(s) - LEN + w);
                if constexpr (LOAD) *p = stw[(size_t)w * sstride + lane];
                else stw[(size_t)w * sstride + lane] = *p;
            }
            StateIO<K, s + 1, LOAD>::run(sm, st, sstride, lane0, nl, tid);
        }
    }
};

// FM = false: x, y lane-major.  FM = true: frame-major x[t][lane], y[t][lane][2^K]: the
// input tile is prefetched element-wise and the staged output rows leave as 16-byte (x2: 8-byte) pieces
// (8 lanes x 4 pieces of one frame per warp store = up to 512 contiguous bytes of HBM).
template <int K, bool FM>
__global__ void __launch_bounds__(NT, 4)
hbf_int_fast_kernel(float *st, const float *x, float *y, size_t n_in, size_t ntiles, size_t lanes, size_t sstride) {
    constexpr int TI = ti(K);
    constexpr int NV = FM ? NL * TI : NL * TI / 4;  // loads per input tile (floats if FM, float4 else)
    constexpr int NVT = (NV + NT - 1) / NT;         // ... per thread
    extern __shared__ __align__(128) float sm[];
    const int tid = threadIdx.x;
    const size_t lane0 = (size_t)blockIdx.x * NL;
    const int nl = (int)((lanes - lane0) < (size_t)NL ? (lanes - lane0) : (size_t)NL);
    const size_t n_out = n_in << K;  // row stride of y in floats

    for (int i = tid; i < smem_floats(K); i += NT) sm[i] = 0.f;
    __syncthreads();
    StateIO<K, 0, true>::run(sm, st, sstride, lane0, nl, tid);

    // input prefetch (the input is 1/2^K of the traffic): float4 v = tid + j*NT of the tile
    float4 nxt[NVT];  // FM uses .x only
    auto fetch = [&](size_t tile) {
#pragma unroll
        for (int j = 0; j < NVT; j++) {
            const int v = tid + j * NT;
            if constexpr (FM) {
                const int plane = v % NL, t = v / NL;  // lane-fastest: 8 lanes of a frame are contiguous
                nxt[j].x = (v < NV && plane < nl) ? x[(tile * TI + t) * lanes + lane0 + plane] : 0.f;
            } else {
                const int plane = v / (TI / 4), pvec = v % (TI / 4);
                nxt[j] = (v < NV && plane < nl)
                             ? *reinterpret_cast<const float4 *>(x + (lane0 + plane) * n_in + tile * TI + 4 * pvec)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    if (ntiles) fetch(0);

    for (size_t i = 0; i < ntiles; i++) {
        const int ob = (int)(i & 1);
#pragma unroll
        for (int j = 0; j < NVT; j++) {
            const int v = tid + j * NT;
            if constexpr (FM) {
                const int plane = v % NL, t = v / NL;
                if (v < NV) sm[off_u(K, 0) + plane * pitch(K, 0) + hist(0) + t] = nxt[j].x;
            } else {
                const int plane = v / (TI / 4), pvec = v % (TI / 4);
                if (v < NV) *reinterpret_cast<float4 *>(sm + off_u(K, 0) + plane * pitch(K, 0) + hist(0) + 4 * pvec) = nxt[j];
            }
        }
        if (i + 1 < ntiles) fetch(i + 1);
        // rows K-1 of the previous tile: last read in its final phase, next written by stage K-2
        if constexpr (K >= 2) {
            if (i > 0) carry_rows<K, K - 1>(sm, tid >> 5, tid & 31);
        }
        // the staging buffer about to be refilled must have been drained by its bulk stores
        // (bulk async-groups are per thread: every issuing thread waits for its own)
        if constexpr (!FM) {
            if (tid < nl) tma_wait_read<1>();
        }
        __syncthreads();
        if constexpr (K >= 2) { StageRun<K, 0>::run(sm, tid, ob); __syncthreads(); }
        if constexpr (K >= 3) { StageRun<K, 1>::run(sm, tid, ob); __syncthreads(); }
        if constexpr (K >= 4) { StageRun<K, 2>::run(sm, tid, ob); __syncthreads(); }
        if constexpr (K >= 5) { StageRun<K, 3>::run(sm, tid, ob); __syncthreads(); }
        StageRun<K, K - 1>::run(sm, tid, ob);  // -> staging (and carries rows K-2)
        if constexpr (!FM) fence_async_smem();  // writers make the staging rows visible to the async proxy
        __syncthreads();
        if constexpr (FM) {
            constexpr int R = 1 << K;
            constexpr int PF_ = R >= 4 ? 4 : 2;  // floats per piece: 16 bytes, or the 8-byte frame of x2
            const float *stg = sm + off_out(K) + ob * NL * OUT_PITCH;
            for (int c = tid; c < NL * TOUT / PF_; c += NT) {
                const int l = c % NL, q = c / NL;  // piece q = output samples PF_*q .. of lane l's tile
                if (l < nl) {
                    float *dst = y + ((i * TI + (PF_ * q) / R) * lanes + lane0 + l) * R + (PF_ * q) % R;
                    if constexpr (PF_ == 4) {
                        *reinterpret_cast<float4 *>(dst) = lds128v(stg + l * OUT_PITCH + PF_ * q);
                    } else {
                        *reinterpret_cast<float2 *>(dst) = *reinterpret_cast<const float2 *>(stg + l * OUT_PITCH + PF_ * q);
                    }
                }
            }
        } else if (tid < nl) {
            bulk_store_1d(y + (lane0 + tid) * n_out + i * TOUT, smem_u32(sm + off_out(K) + (ob * NL + tid) * OUT_PITCH),
                          TOUT * 4);
            tma_commit();
        }
        if constexpr (K == 1) {  // rows 0 are rewritten at the top of the next tile: carry them now
            carry_rows<K, 0>(sm, tid >> 5, tid & 31);
            __syncthreads();
        }
    }
    if constexpr (K >= 2) {
        carry_rows<K, K - 1>(sm, tid >> 5, tid & 31);
        __syncthreads();
    }
    if constexpr (!FM) {
        if (tid < nl) tma_wait_read<0>();
    }
    StateIO<K, 0, false>::run(sm, st, sstride, lane0, nl, tid);
}

template <int K, bool FM>
static int launch(idsp_ctx *ctx, float *st, const float *x, float *y, size_t n_in, size_t ntiles, size_t lanes,
                  size_t sstride) {
    auto kern = hbf_int_fast_kernel<K, FM>;
    IDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(K)));
    unsigned grid = (unsigned)((lanes + NL - 1) / NL);
    kern<<<grid, NT, smem_bytes(K), ctx->stream>>>(st, x, y, n_in, ntiles, lanes, sstride);
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

}  // namespace hfi

// Runs the tiled kernel over the first (n_in / TI) * TI input frames of every lane; *done =
// frames covered (the generic kernel finishes the tail), or IDSP_HBF_FAST_NOT_APPLICABLE.
static int hbf_int_fast_try(idsp_ctx *ctx, int k, float *state, const float *x, float *y, size_t n_in,
                            size_t lanes, size_t sstride, int layout, size_t *done) {
    *done = 0;
    const bool fm = layout == IDSP_FRAME_MAJOR;
    // frame-major x32 measured faster on the generic thread-per-lane kernel (714 vs 560 GSa/s)
    if (ctx->policy == 1 || (fm && k > 4 && ctx->policy != 2)) return IDSP_HBF_FAST_NOT_APPLICABLE;
    const size_t TI = (size_t)hfi::TOUT >> k;
    const size_t ntiles = n_in / TI;
    const bool ok = ntiles >= 1 && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0 && (fm || (n_in % 4) == 0);
    if (!ok) return IDSP_HBF_FAST_NOT_APPLICABLE;
    int r;
    if (fm) {
        switch (k) {
            case 1: r = hfi::launch<1, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
            case 2: r = hfi::launch<2, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
            case 3: r = hfi::launch<3, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
            case 4: r = hfi::launch<4, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
            default: r = hfi::launch<5, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
        }
    } else {
        switch (k) {
            case 1: r = hfi::launch<1, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
            case 2: r = hfi::launch<2, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
            case 3: r = hfi::launch<3, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
            case 4: r = hfi::launch<4, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
            default: r = hfi::launch<5, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break;
        }
    }
    if (r == IDSP_OK) *done = ntiles * TI;
    return r;
}

}  // namespace idsp
