// hbf_fast_scalar.cuh -- shared-memory tiled HBF /2^K decimation cascade (f32, lane-major),
// scalar FP32 variant (the default; hbf_fast.cuh holds the packed f32x2 variant, both bit-exact).
//
// The FIR stages are time-parallel, so unlike the biquad kernels a lane is not tied to one
// thread.  A CTA owns NL = 8 lanes for the whole call and walks the time axis in tiles of
// TT = 512 input samples per lane (4 CTAs of 128 threads per SM, each its own barrier domain):
//
//   HBM --TMA 1-D bulk copy per lane row (2 KB), mbarrier complete_tx--> raw ring (S = 2)
//   stage 0 : reads the interleaved raw stream, writes de-interleaved E/O rows of stage 1
//   stage s : reads E_s / O_s, writes E_{s+1} / O_{s+1}          (all in shared memory)
//   stage K-1 : writes the decimated output straight to HBM
//
// Work item = (lane, R consecutive outputs of one stage); items are spread over the CTA's
// threads lane-fastest, so a quarter-warp touches 8 different rows whose pitch is 4*odd
// floats -> every LDS.128 / STS.128 is bank-conflict free.  A thread loads its whole window
// into registers with static indices (no shifting delay line) and evaluates R outputs in
// exactly the reference's order:
//   acc = ((w[2M-1]+w[0])*c0) + ((w[2M-2]+w[1])*c1) + ...  then  + even sample
// (src/hbf.rs:46-68, :178-181), each op individually rounded (-fmad=false) -> bit-exact.
// Each row keeps the history the next tile needs ([hist | tile]); after a tile the tails are
// moved to the heads (the reference's copy_within, src/hbf.rs:183-184).  The ABI state
// (even/odd history per stage) is scattered into those heads at entry and gathered back at
// exit, so calls can be chained like block() calls on the reference.
// Tile-shape measurements (262144 lanes x 65536 inputs, /16): NL16/NT256/R0=8 826 GSa/s,
// NL16/NT256/R0=16 846, NL8/NT64 873, NL8/NT128/R0=16 901 (profiles/r1_hbf_variants.log).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "hbf_stages.cuh"
#include "tma_kernels.cuh"

#ifndef IDSP_HBF_FAST_NOT_APPLICABLE
#define IDSP_HBF_FAST_NOT_APPLICABLE 12346
#endif

namespace idsp {
namespace hfs {

#ifndef HFS_NL
#define HFS_NL 8
#endif
constexpr int NL = HFS_NL;  // lanes per CTA (multiple of 8)
#ifndef HFS_NT
#define HFS_NT 128
#endif
constexpr int NT = HFS_NT;  // threads per CTA
constexpr int TT = 512;   // raw input samples per lane per tile
constexpr int S = 2;      // raw ring depth

__host__ __device__ constexpr int up4(int v) { return (v + 3) & ~3; }
// pitch in floats: multiple of 4 with pitch/4 odd (conflict-free 16-byte accesses over 8 rows)
__host__ __device__ constexpr int oddpitch(int v) { return (up4(v) / 4) % 2 ? up4(v) : up4(v) + 4; }
__host__ __device__ constexpr int st_m(int K, int s) { return hbf_m(K - 1 - s); }
__host__ __device__ constexpr int st_n(int s) { return TT >> (s + 1); }  // outputs per lane per tile
#ifndef HFS_MINB
#define HFS_MINB 4
#endif
#ifndef HFS_R0
#define HFS_R0 16
#endif
__host__ __device__ constexpr int st_r(int s) { return s == 0 ? HFS_R0 : (st_n(s) / 8 >= 8 ? 8 : (st_n(s) / 8 >= 4 ? st_n(s) / 8 : 4)); }
__host__ __device__ constexpr int raw_h(int K) { return up4(4 * st_m(K, 0) - 2); }
__host__ __device__ constexpr int raw_pitch(int K) { return oddpitch(raw_h(K) + TT); }
__host__ __device__ constexpr int he(int K, int s) { return up4(st_m(K, s) - 1); }
__host__ __device__ constexpr int ho(int K, int s) { return up4(2 * st_m(K, s) - 1); }
__host__ __device__ constexpr int pe(int K, int s) { return oddpitch(he(K, s) + st_n(s)); }
__host__ __device__ constexpr int po(int K, int s) { return oddpitch(ho(K, s) + st_n(s)); }
// float offsets inside dynamic shared memory
__host__ __device__ constexpr int off_e(int K, int s) {
    int o = S * NL * raw_pitch(K);
    for (int i = 1; i < s; i++) o += NL * (pe(K, i) + po(K, i));
    return o;
}
__host__ __device__ constexpr int off_o(int K, int s) { return off_e(K, s) + NL * pe(K, s); }
__host__ __device__ constexpr int smem_floats(int K) { return off_e(K, K); }
__host__ __device__ constexpr size_t smem_bytes(int K) { return (size_t)smem_floats(K) * 4 + S * 8; }
// ABI state word offset of stage s (highest-rate stage first): sum of 3M-2
__host__ __device__ constexpr int st_word(int K, int s) {
    int w = 0;
    for (int i = 0; i < s; i++) w += 3 * st_m(K, i) - 2;
    return w;
}

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// 128-bit shared loads as explicit PTX: the compiler must not split them into scalar LDS
// (only some components of a window are used, and scalar loads over rows of pitch 4*odd
// floats would be 4-way bank conflicted).
__device__ __forceinline__ float4 lds128(const float *p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(smem_u32(p)));
    return v;
}

// One item of the raw (interleaved) stage: outputs p0 .. p0+R-1 of lane row `row`
// (row[0..HR) = history, row[HR..] = tile).  Stream sample u[k] (k relative to the tile
// start) sits at row[HR + k]; the window starts at row[2*p0] (16-byte aligned).
template <int TI, int R> struct RawItem {
    static constexpr int M = HbfTaps<TI>::M;
    static constexpr int HR = up4(4 * M - 2);
    static constexpr int W = HR + 2 * R;
    __device__ __forceinline__ static void run(const float *row, int p0, float (&y)[R]) {
        float w[W];
        const float *src = row + 2 * p0;
#pragma unroll
        for (int j = 0; j < W / 4; j++) {
            float4 v = lds128(src + 4 * j);
            w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < R; q++) {
            float acc = (w[2 * q + 1 + HR] + w[2 * q - 4 * M + 3 + HR]) * HbfTaps<TI>::c(0);
#pragma unroll
            for (int i = 1; i < M; i++)
                acc = acc + (w[2 * q - 2 * i + 1 + HR] + w[2 * q + 2 * i - 4 * M + 3 + HR]) * HbfTaps<TI>::c(i);
            y[q] = acc + w[2 * q - 2 * M + 2 + HR];
        }
    }
};

// One item of a de-interleaved stage: erow = [HE hist | n new], orow = [HO hist | n new].
template <int TI, int R> struct SplitItem {
    static constexpr int M = HbfTaps<TI>::M;
    static constexpr int LEN = 2 * M - 1;
    static constexpr int HE = up4(M - 1), HO = up4(LEN);
    static constexpr int RE = HE - (M - 1), RO = HO - LEN;
    static constexpr int WO = up4(RO + R + 2 * M - 1), WE = up4(RE + R);
    __device__ __forceinline__ static void run(const float *erow, const float *orow, int p0, float (&y)[R]) {
        float wo[WO], we[WE];
#pragma unroll
        for (int j = 0; j < WO / 4; j++) {
            float4 v = lds128(orow + p0 + 4 * j);
            wo[4 * j] = v.x; wo[4 * j + 1] = v.y; wo[4 * j + 2] = v.z; wo[4 * j + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < WE / 4; j++) {
            float4 v = lds128(erow + p0 + 4 * j);
            we[4 * j] = v.x; we[4 * j + 1] = v.y; we[4 * j + 2] = v.z; we[4 * j + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < R; q++) {
            float acc = (wo[RO + q + 2 * M - 1] + wo[RO + q]) * HbfTaps<TI>::c(0);
#pragma unroll
            for (int i = 1; i < M; i++)
                acc = acc + (wo[RO + q + 2 * M - 1 - i] + wo[RO + q + i]) * HbfTaps<TI>::c(i);
            y[q] = acc + we[RE + q];
        }
    }
};

// scatter R consecutive outputs (p0 multiple of R, R in {4,8}) into the next stage's E/O rows
template <int R>
__device__ __forceinline__ void put_split(float *erow_new, float *orow_new, int p0, const float (&y)[R]) {
    // erow_new / orow_new already point at the first NEW element (past the history)
    if constexpr (R >= 8) {
#pragma unroll
        for (int j = 0; j < R / 8; j++) {
            reinterpret_cast<float4 *>(erow_new + p0 / 2)[j] = make_float4(y[8 * j], y[8 * j + 2], y[8 * j + 4], y[8 * j + 6]);
            reinterpret_cast<float4 *>(orow_new + p0 / 2)[j] = make_float4(y[8 * j + 1], y[8 * j + 3], y[8 * j + 5], y[8 * j + 7]);
        }
    } else {
        *reinterpret_cast<float2 *>(erow_new + p0 / 2) = make_float2(y[0], y[2]);
        *reinterpret_cast<float2 *>(orow_new + p0 / 2) = make_float2(y[1], y[3]);
    }
}

template <int K, int s> struct StageRun {
    // runs stage s (1 <= s <= K-1) for one tile
    __device__ __forceinline__ static void run(float *sm, int tid, int nl, float *y, size_t ystride,
                                               size_t yoff, size_t lane0) {
        constexpr int TI = K - 1 - s;
        constexpr int R = st_r(s);
        constexpr int ITEMS = NL * st_n(s) / R;
        const float *E = sm + off_e(K, s);
        const float *O = sm + off_o(K, s);
        for (int idx = tid; idx < ITEMS; idx += NT) {
            const int lane = idx % NL, p0 = (idx / NL) * R;
            float out[R];
            SplitItem<TI, R>::run(E + lane * pe(K, s), O + lane * po(K, s), p0, out);
            if constexpr (s == K - 1) {
                if (lane < nl) {
                    float *dst = y + (lane0 + lane) * ystride + yoff + p0;
                    if ((((uintptr_t)dst) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < R / 4; j++)
                            reinterpret_cast<float4 *>(dst)[j] =
                                make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < R; j++) dst[j] = out[j];
                    }
                }
            } else {
                float *En = sm + off_e(K, s + 1) + lane * pe(K, s + 1) + he(K, s + 1);
                float *On = sm + off_o(K, s + 1) + lane * po(K, s + 1) + ho(K, s + 1);
                put_split<R>(En, On, p0, out);
            }
        }
    }
};

// move the tail of a [hist | n new] row to its head; one thread per row, through registers
template <int H, int N> __device__ __forceinline__ void carry_row(float *row) {
    float t[H];
#pragma unroll
    for (int j = 0; j < H / 4; j++) {
        float4 v = lds128(row + N + 4 * j);
        t[4 * j] = v.x; t[4 * j + 1] = v.y; t[4 * j + 2] = v.z; t[4 * j + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < H / 4; j++)
        reinterpret_cast<float4 *>(row)[j] = make_float4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
}

template <int K, int s> struct Carry {
    __device__ __forceinline__ static void run(float *sm, int job, int lane) {
        if constexpr (s < K) {
            if (job == 2 * (s - 1)) carry_row<he(K, s), st_n(s)>(sm + off_e(K, s) + lane * pe(K, s));
            else if (job == 2 * (s - 1) + 1) carry_row<ho(K, s), st_n(s)>(sm + off_o(K, s) + lane * po(K, s));
            else Carry<K, s + 1>::run(sm, job, lane);
        }
    }
};

// ABI state <-> shared-memory histories (see header comment of include/idsp_b200.h)
template <int K, int s, bool LOAD> struct StateIO {
    __device__ __forceinline__ static void run(float *sm, float *st, size_t sstride, size_t lane0, int nl,
                                               int tid, int rawbuf) {
        if constexpr (s < K) {
            constexpr int M = st_m(K, s);
            constexpr int LEN = 2 * M - 1;
            constexpr int WORDS = 3 * M - 2;
            float *stw = st + (size_t)st_word(K, s) * sstride + lane0;
            for (int idx = tid; idx < WORDS * NL; idx += NT) {
                const int lane = idx % NL, w = idx / NL;
                if (lane >= nl) continue;
                float *p;
                if constexpr (s == 0) {
                    constexpr int HR = raw_h(K);
                    float *row = sm + (rawbuf * NL + lane) * raw_pitch(K);
                    p = w < M - 1 ? row + (HR - 2 * M + 2 + 2 * w) : row + (HR - 4 * M + 3 + 2 * (w - (M - 1)));
                } else {
                    p = w < M - 1 ? sm + off_e(K, s) + lane * pe(K, s) + (he(K, s) - (M - 1) + w)
                                  : sm + off_o(K, s) + lane * po(K, s) + (ho(K, s) - LEN + (w - (M - 1)));
                }
                if constexpr (LOAD) *p = stw[(size_t)w * sstride + lane];
                else stw[(size_t)w * sstride + lane] = *p;
            }
            StateIO<K, s + 1, LOAD>::run(sm, st, sstride, lane0, nl, tid, rawbuf);
        }
    }
};

template <int K>
__global__ void __launch_bounds__(NT, HFS_MINB)
hbf_dec_fast_kernel(float *st, const float *x, float *y, size_t n_out, size_t ntiles, size_t lanes,
                    size_t sstride) {
    constexpr int TI0 = K - 1;
    constexpr int R0 = st_r(0);
    constexpr int HR = raw_h(K);
    constexpr int PR = raw_pitch(K);
    constexpr int TO = TT >> K;
    extern __shared__ __align__(128) float sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + smem_floats(K));
    const int tid = threadIdx.x;
    const size_t lane0 = (size_t)blockIdx.x * NL;
    const int nl = (int)((lanes - lane0) < (size_t)NL ? (lanes - lane0) : (size_t)NL);
    const size_t n_in = n_out << K;  // row stride of x in floats

    // zero everything once (unused history slots / absent lanes must hold finite garbage-free data)
    for (int i = tid; i < smem_floats(K); i += NT) sm[i] = 0.f;
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < S; b++) mbar_init(smem_u32(&bars[b]), 1);
        mbar_fence_init();
    }
    __syncthreads();
    StateIO<K, 0, true>::run(sm, st, sstride, lane0, nl, tid, 0);
    // generic-proxy writes above (zero fill) precede async-proxy (TMA) writes to the same rows
    fence_async_smem();
    __syncthreads();

    auto issue = [&](size_t tile) {  // executed by warp 0
        const int b = (int)(tile % S);
        const uint32_t bar = smem_u32(&bars[b]);
        if ((tid & 31) == 0) mbar_expect_tx(bar, (uint32_t)(nl * TT * 4));
        __syncwarp();
        if (tid < nl)
            bulk_load_1d(smem_u32(sm + (b * NL + tid) * PR + HR), x + (lane0 + tid) * n_in + tile * TT, TT * 4, bar);
    };
    if (tid < 32) {
#pragma unroll
        for (int b = 0; b < S; b++)
            if ((size_t)b < ntiles) issue(b);
    }

    for (size_t i = 0; i < ntiles; i++) {
        const int b = (int)(i % S);
        mbar_wait(smem_u32(&bars[b]), (uint32_t)((i / S) & 1));
        // ---- stage 0: raw interleaved -> E_1 / O_1 (or -> y when K == 1)
        {
            const float *raw = sm + b * NL * PR;
            constexpr int ITEMS = NL * st_n(0) / R0;
            for (int idx = tid; idx < ITEMS; idx += NT) {
                const int lane = idx % NL, p0 = (idx / NL) * R0;
                float out[R0];
                RawItem<TI0, R0>::run(raw + lane * PR, p0, out);
                if constexpr (K == 1) {
                    if (lane < nl) {
                        float *dst = y + (lane0 + lane) * n_out + i * TO + p0;
                        if ((((uintptr_t)dst) & 15) == 0) {
#pragma unroll
                            for (int j = 0; j < R0 / 4; j++)
                                reinterpret_cast<float4 *>(dst)[j] =
                                    make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < R0; j++) dst[j] = out[j];
                        }
                    }
                } else {
                    float *En = sm + off_e(K, 1) + lane * pe(K, 1) + he(K, 1);
                    float *On = sm + off_o(K, 1) + lane * po(K, 1) + ho(K, 1);
                    put_split<R0>(En, On, p0, out);
                }
            }
        }
        __syncthreads();
        // ---- raw history: tail of buffer b -> head of the next buffer, then refill buffer b
        if (tid < 32) {
            if (tid < NL) {
                float *src = sm + (b * NL + tid) * PR;
                float *dst = sm + (((b + 1) % S) * NL + tid) * PR;
                float t[HR];
#pragma unroll
                for (int j = 0; j < HR / 4; j++) {
                    float4 v = lds128(src + TT + 4 * j);
                    t[4 * j] = v.x; t[4 * j + 1] = v.y; t[4 * j + 2] = v.z; t[4 * j + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < HR / 4; j++)
                    reinterpret_cast<float4 *>(dst)[j] = make_float4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
            }
            __syncwarp();
            if (i + S < ntiles) issue(i + S);
        }
        // ---- stages 1 .. K-1
        if constexpr (K >= 2) { StageRun<K, 1>::run(sm, tid, nl, y, n_out, i * TO, lane0); __syncthreads(); }
        if constexpr (K >= 3) { StageRun<K, 2>::run(sm, tid, nl, y, n_out, i * TO, lane0); __syncthreads(); }
        if constexpr (K >= 4) { StageRun<K, 3>::run(sm, tid, nl, y, n_out, i * TO, lane0); __syncthreads(); }
        if constexpr (K >= 5) { StageRun<K, 4>::run(sm, tid, nl, y, n_out, i * TO, lane0); __syncthreads(); }
        // ---- carry the E/O histories (one thread per row)
        if constexpr (K >= 2) {
            for (int idx = tid; idx < 2 * (K - 1) * NL; idx += NT) Carry<K, 1>::run(sm, idx / NL, idx % NL);
        }
        // also orders warp 0's raw-history copy before the next tile's stage 0
        __syncthreads();
    }
    // raw history of the stream now sits at the head of buffer (ntiles % S)
    __syncthreads();
    StateIO<K, 0, false>::run(sm, st, sstride, lane0, nl, tid, (int)(ntiles % S));
}

template <int K>
static int launch(idsp_ctx *ctx, float *st, const float *x, float *y, size_t n_out, size_t ntiles,
                  size_t lanes, size_t sstride) {
    auto kern = hbf_dec_fast_kernel<K>;
    IDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(K)));
    unsigned grid = (unsigned)((lanes + NL - 1) / NL);
    kern<<<grid, NT, smem_bytes(K), ctx->stream>>>(st, x, y, n_out, ntiles, lanes, sstride);
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

}  // namespace hfs

// Runs the tiled kernel over the first (n_out / TO) * TO frames of every lane.  Returns
// the number of frames it covered in *done (the caller finishes the tail with the
// generic kernel), or IDSP_HBF_FAST_NOT_APPLICABLE.
static int hbf_dec_fast_try_scalar(idsp_ctx *ctx, int k, float *state, const float *x, float *y, size_t n_out,
                            size_t lanes, size_t sstride, int layout, size_t *done) {
    *done = 0;
    if (ctx->policy == 1 || layout != IDSP_LANE_MAJOR) return IDSP_HBF_FAST_NOT_APPLICABLE;
    const size_t TO = (size_t)hfs::TT >> k;
    const size_t ntiles = n_out / TO;
    const bool ok = ntiles >= 1 && (((uintptr_t)x) & 15) == 0 && ((n_out << k) % 4) == 0;
    if (!ok) return IDSP_HBF_FAST_NOT_APPLICABLE;
    int r;
    switch (k) {
        case 1: r = hfs::launch<1>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
        case 2: r = hfs::launch<2>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
        case 3: r = hfs::launch<3>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
        case 4: r = hfs::launch<4>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
        default: r = hfs::launch<5>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
    }
    if (r == IDSP_OK) *done = ntiles * TO;
    return r;
}

}  // namespace idsp
