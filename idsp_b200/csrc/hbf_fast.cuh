// hbf_fast.cuh -- shared-memory tiled HBF /2^K decimation cascade (f32, lane-major).
//
// The FIR stages are time-parallel, so unlike the biquad kernels a lane is not tied
// to one thread.  A CTA owns NL = 16 lanes (8 lane PAIRS) for the whole call and walks
// the time axis in tiles of TT = 512 input samples per lane:
//
//   HBM --TMA 1-D bulk copy per lane row (2 KB), mbarrier complete_tx--> raw ring (S = 2)
//   stage 0 : reads the interleaved raw stream of both lanes of a pair (scalar FP32) and
//             writes de-interleaved even/odd rows for stage 1 with the two lanes of the
//             pair PACKED side by side: element e of a pair row = (lane A, lane B)
//   stage s : reads E_s / O_s, writes E_{s+1} / O_{s+1}, all in shared memory, all
//             arithmetic as packed add.rn.f32x2 / mul.rn.f32x2 (FADD2/FMUL2: two IEEE
//             round-to-nearest results per instruction, bit-identical to scalar ops;
//             measured 2x the scalar FP32 rate on B200, tools/ubench.cu)
//   stage K-1 : unpacks and writes the decimated output straight to HBM
//
// Work item = (lane pair, R consecutive outputs of one stage); items are spread over
// the CTA's threads pair-fastest, so a quarter-warp touches 8 different rows whose
// pitch is 4*odd floats -> every LDS.128 / STS.128 is bank-conflict free.  A thread
// loads its whole window into registers with static indices (no shifting delay line)
// and evaluates R outputs in exactly the reference's order:
//   acc = ((w[2M-1]+w[0])*c0) + ((w[2M-2]+w[1])*c1) + ...  then  + even sample
// (src/hbf.rs:46-68, :178-181), each op individually rounded -> bit-exact.
// Each row keeps the history the next tile needs ([hist | tile]); after a tile the
// tails are moved to the heads (the reference's copy_within, src/hbf.rs:183-184).
// The ABI state (even/odd history per stage) is scattered into those heads at entry and
// gathered back at exit, so calls can be chained like block() calls on the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "hbf_stages.cuh"
#include "tma_kernels.cuh"

#define IDSP_HBF_FAST_NOT_APPLICABLE 12346

namespace idsp {
namespace hf {

constexpr int NL = 16;      // lanes per CTA
constexpr int NP = NL / 2;  // lane pairs per CTA
#ifndef HF_NT
#define HF_NT 128
#endif
constexpr int NT = HF_NT;  // threads per CTA
constexpr int TT = 512;    // raw input samples per lane per tile
constexpr int S = 2;       // raw ring depth

__host__ __device__ constexpr int up2(int v) { return (v + 1) & ~1; }
__host__ __device__ constexpr int up4(int v) { return (v + 3) & ~3; }
// pitch in floats: multiple of 4 with pitch/4 odd (conflict-free 16-byte accesses over 8 rows)
__host__ __device__ constexpr int oddpitch(int v) { return (up4(v) / 4) % 2 ? up4(v) : up4(v) + 4; }
__host__ __device__ constexpr int st_m(int K, int s) { return hbf_m(K - 1 - s); }
__host__ __device__ constexpr int st_n(int s) { return TT >> (s + 1); }  // outputs per lane per tile
// outputs per item: stage 0 (scalar, per lane) 8; packed stages n/16 clamped to [2, 8]
__host__ __device__ constexpr int st_r(int s) {
    return s == 0 ? 8 : (st_n(s) / 16 >= 8 ? 8 : (st_n(s) / 16 >= 2 ? st_n(s) / 16 : 2));
}
__host__ __device__ constexpr int raw_h(int K) { return up4(4 * st_m(K, 0) - 2); }
__host__ __device__ constexpr int raw_pitch(int K) { return oddpitch(raw_h(K) + TT); }
// packed rows (s >= 1): history in ELEMENTS (one element = 2 floats = both lanes of a pair)
__host__ __device__ constexpr int he(int K, int s) { return up2(st_m(K, s) - 1); }
__host__ __device__ constexpr int ho(int K, int s) { return up2(2 * st_m(K, s) - 1); }
__host__ __device__ constexpr int pe(int K, int s) { return oddpitch(2 * (he(K, s) + st_n(s))); }  // floats
__host__ __device__ constexpr int po(int K, int s) { return oddpitch(2 * (ho(K, s) + st_n(s))); }
// float offsets inside dynamic shared memory
__host__ __device__ constexpr int off_e(int K, int s) {
    int o = S * NL * raw_pitch(K);
    for (int i = 1; i < s; i++) o += NP * (pe(K, i) + po(K, i));
    return o;
}
__host__ __device__ constexpr int off_o(int K, int s) { return off_e(K, s) + NP * pe(K, s); }
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

// 128-bit shared loads as explicit *volatile* PTX: ptxas narrows a plain `ld.shared.v4`
// whose components are partly unused (windows over the interleaved stream) to scalar
// LDS / LDS.64, and those are 4-8 way bank conflicted over rows of pitch 4*odd floats
// (seen in profiles/r1_hbf_*: 147 M excess wavefronts).  `ld.volatile` keeps LDS.128.
__device__ __forceinline__ float4 lds128(const float *p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(smem_u32(p)));
    return v;
}
typedef unsigned long long f2;  // two packed f32: lo = lane A, hi = lane B
__device__ __forceinline__ void lds_2f2(const float *p, f2 &a, f2 &b) {
    asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void sts_2f2(float *p, f2 a, f2 b) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(smem_u32(p)), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void sts_f2(float *p, f2 a) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(smem_u32(p)), "l"(a) : "memory");
}
__device__ __forceinline__ f2 pk(float a, float b) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpk(f2 v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// Packed multiply.  ptxas (12.9) contracts `mul.rn.f32x2` + `add.rn.f32x2` into one FFMA2
// even with --fmad false (it does not for scalar .rn ops), which would change the
// rounding.  So the product is written as fma(a, b, nz) with nz = (-0.0, -0.0) supplied
// as an opaque kernel parameter: x*y + (-0.0) rounds exactly like x*y (sign of zero
// included), costs the same single instruction, and cannot be merged with the add that
// follows.  tests/test_gpu_hbf.py compares bit patterns against the scalar CPU restatement.
__device__ __forceinline__ f2 mul2(f2 a, f2 b, f2 nz) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}

// One lane of the raw (interleaved) stage: outputs p0 .. p0+R-1 of lane row `row`
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

// One item of a packed stage: erow = [HE hist | n new] elements, orow = [HO hist | n new],
// element = (lane A, lane B).  p0 (first output, even) is also the element offset of the
// window inside the rows.
template <int TI, int R> struct PackedItem {
    static constexpr int M = HbfTaps<TI>::M;
    static constexpr int LEN = 2 * M - 1;
    static constexpr int HE = up2(M - 1), HO = up2(LEN);
    static constexpr int RE = HE - (M - 1), RO = HO - LEN;
    static constexpr int WO = up2(RO + R + 2 * M - 1), WE = up2(RE + R);
    __device__ __forceinline__ static void run(const float *erow, const float *orow, int p0, f2 nz,
                                               f2 (&y)[R]) {
        f2 wo[WO], we[WE];
#pragma unroll
        for (int j = 0; j < WO / 2; j++) lds_2f2(orow + 2 * p0 + 4 * j, wo[2 * j], wo[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < WE / 2; j++) lds_2f2(erow + 2 * p0 + 4 * j, we[2 * j], we[2 * j + 1]);
#pragma unroll
        for (int q = 0; q < R; q++) {
            const float c0 = HbfTaps<TI>::c(0);
            f2 acc = mul2(add2(wo[RO + q + 2 * M - 1], wo[RO + q]), pk(c0, c0), nz);
#pragma unroll
            for (int i = 1; i < M; i++) {
                const float ci = HbfTaps<TI>::c(i);
                acc = add2(acc, mul2(add2(wo[RO + q + 2 * M - 1 - i], wo[RO + q + i]), pk(ci, ci), nz));
            }
            y[q] = add2(acc, we[RE + q]);
        }
    }
};

// scatter R consecutive packed outputs (p0 multiple of R) into the next stage's E/O rows
template <int R>
__device__ __forceinline__ void put_packed(float *e_new, float *o_new, int p0, const f2 (&y)[R]) {
    // e_new / o_new point at the first NEW element (past the history); output p goes to
    // element p/2 of the even (p even) or odd (p odd) row, i.e. float offset p0 + ...
    if constexpr (R == 8) {
        sts_2f2(e_new + p0, y[0], y[2]);
        sts_2f2(e_new + p0 + 4, y[4], y[6]);
        sts_2f2(o_new + p0, y[1], y[3]);
        sts_2f2(o_new + p0 + 4, y[5], y[7]);
    } else if constexpr (R == 4) {
        sts_2f2(e_new + p0, y[0], y[2]);
        sts_2f2(o_new + p0, y[1], y[3]);
    } else {
        sts_f2(e_new + p0, y[0]);
        sts_f2(o_new + p0, y[1]);
    }
}

template <int R> __device__ __forceinline__ void store_out(float *dst, const float (&v)[R]) {
    if (R % 4 == 0 && (((uintptr_t)dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < R / 4; j++)
            reinterpret_cast<float4 *>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else if (R % 2 == 0 && (((uintptr_t)dst) & 7) == 0) {
#pragma unroll
        for (int j = 0; j < R / 2; j++) reinterpret_cast<float2 *>(dst)[j] = make_float2(v[2 * j], v[2 * j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < R; j++) dst[j] = v[j];
    }
}

template <int K, int s> struct StageRun {
    // runs packed stage s (1 <= s <= K-1) for one tile
    __device__ __forceinline__ static void run(float *sm, int tid, int nl, float *y, size_t ystride,
                                               size_t yoff, size_t lane0, f2 nz) {
        constexpr int TI = K - 1 - s;
        constexpr int R = st_r(s);
        constexpr int ITEMS = NP * st_n(s) / R;
        const float *E = sm + off_e(K, s);
        const float *O = sm + off_o(K, s);
        for (int idx = tid; idx < ITEMS; idx += NT) {
            const int pr = idx % NP, p0 = (idx / NP) * R;
            f2 out[R];
            PackedItem<TI, R>::run(E + pr * pe(K, s), O + pr * po(K, s), p0, nz, out);
            if constexpr (s == K - 1) {
                float a[R], b[R];
#pragma unroll
                for (int q = 0; q < R; q++) unpk(out[q], a[q], b[q]);
                if (2 * pr < nl) store_out<R>(y + (lane0 + 2 * pr) * ystride + yoff + p0, a);
                if (2 * pr + 1 < nl) store_out<R>(y + (lane0 + 2 * pr + 1) * ystride + yoff + p0, b);
            } else {
                float *En = sm + off_e(K, s + 1) + pr * pe(K, s + 1) + 2 * he(K, s + 1);
                float *On = sm + off_o(K, s + 1) + pr * po(K, s + 1) + 2 * ho(K, s + 1);
                put_packed<R>(En, On, p0, out);
            }
        }
    }
};

// move the tail of a [H hist | N new] row (in floats) to its head; one thread per row
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
    __device__ __forceinline__ static void run(float *sm, int job, int pr) {
        if constexpr (s < K) {
            if (job == 2 * (s - 1))
                carry_row<2 * he(K, s), 2 * st_n(s)>(sm + off_e(K, s) + pr * pe(K, s));
            else if (job == 2 * (s - 1) + 1)
                carry_row<2 * ho(K, s), 2 * st_n(s)>(sm + off_o(K, s) + pr * po(K, s));
            else
                Carry<K, s + 1>::run(sm, job, pr);
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
                    const int pr = lane >> 1, c = lane & 1;
                    p = w < M - 1 ? sm + off_e(K, s) + pr * pe(K, s) + 2 * (he(K, s) - (M - 1) + w) + c
                                  : sm + off_o(K, s) + pr * po(K, s) + 2 * (ho(K, s) - LEN + (w - (M - 1))) + c;
                }
                if constexpr (LOAD) *p = stw[(size_t)w * sstride + lane];
                else stw[(size_t)w * sstride + lane] = *p;
            }
            StateIO<K, s + 1, LOAD>::run(sm, st, sstride, lane0, nl, tid, rawbuf);
        }
    }
};

template <int K>
__global__ void __launch_bounds__(NT, 2)
hbf_dec_fast_kernel(float *st, const float *x, float *y, size_t n_out, size_t ntiles, size_t lanes,
                    size_t sstride, f2 nz /* (-0.0f, -0.0f), see mul2() */) {
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

    // zero everything once (unused history slots / absent lanes hold zeros, never NaN garbage)
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
        // ---- stage 0: raw interleaved (scalar FP32)
        {
            const float *raw = sm + b * NL * PR;
            if constexpr (K == 1) {  // single stage: straight to HBM, one lane per item
                constexpr int ITEMS = NL * st_n(0) / R0;
                for (int idx = tid; idx < ITEMS; idx += NT) {
                    const int lane = idx % NL, p0 = (idx / NL) * R0;
                    float out[R0];
                    RawItem<TI0, R0>::run(raw + lane * PR, p0, out);
                    if (lane < nl) store_out<R0>(y + (lane0 + lane) * n_out + i * TO + p0, out);
                }
            } else {  // both lanes of a pair -> packed E_1 / O_1
                constexpr int ITEMS = NP * st_n(0) / R0;
                for (int idx = tid; idx < ITEMS; idx += NT) {
                    const int pr = idx % NP, p0 = (idx / NP) * R0;
                    float a[R0], bb[R0];
                    RawItem<TI0, R0>::run(raw + (2 * pr) * PR, p0, a);
                    RawItem<TI0, R0>::run(raw + (2 * pr + 1) * PR, p0, bb);
                    f2 out[R0];
#pragma unroll
                    for (int q = 0; q < R0; q++) out[q] = pk(a[q], bb[q]);
                    float *En = sm + off_e(K, 1) + pr * pe(K, 1) + 2 * he(K, 1);
                    float *On = sm + off_o(K, 1) + pr * po(K, 1) + 2 * ho(K, 1);
                    put_packed<R0>(En, On, p0, out);
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
        // ---- packed stages 1 .. K-1
        if constexpr (K >= 2) { StageRun<K, 1>::run(sm, tid, nl, y, n_out, i * TO, lane0, nz); __syncthreads(); }
        if constexpr (K >= 3) { StageRun<K, 2>::run(sm, tid, nl, y, n_out, i * TO, lane0, nz); __syncthreads(); }
        if constexpr (K >= 4) { StageRun<K, 3>::run(sm, tid, nl, y, n_out, i * TO, lane0, nz); __syncthreads(); }
        if constexpr (K >= 5) { StageRun<K, 4>::run(sm, tid, nl, y, n_out, i * TO, lane0, nz); __syncthreads(); }
        // ---- carry the E/O histories (one thread per row)
        if constexpr (K >= 2) {
            for (int idx = tid; idx < 2 * (K - 1) * NP; idx += NT) Carry<K, 1>::run(sm, idx / NP, idx % NP);
        }
        // also orders warp 0's raw-history copy before the next tile's stage 0
        __syncthreads();
    }
    // raw history of the stream now sits at the head of buffer (ntiles % S)
    StateIO<K, 0, false>::run(sm, st, sstride, lane0, nl, tid, (int)(ntiles % S));
}

template <int K>
static int launch(idsp_ctx *ctx, float *st, const float *x, float *y, size_t n_out, size_t ntiles,
                  size_t lanes, size_t sstride) {
    auto kern = hbf_dec_fast_kernel<K>;
    IDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(K)));
    unsigned grid = (unsigned)((lanes + NL - 1) / NL);
    kern<<<grid, NT, smem_bytes(K), ctx->stream>>>(st, x, y, n_out, ntiles, lanes, sstride,
                                                  0x8000000080000000ull);
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

}  // namespace hf

// Runs the tiled kernel over the first (n_out / TO) * TO frames of every lane.  Returns
// the number of frames it covered in *done (the caller finishes the tail with the
// generic kernel), or IDSP_HBF_FAST_NOT_APPLICABLE.
static int hbf_dec_fast_try(idsp_ctx *ctx, int k, float *state, const float *x, float *y, size_t n_out,
                            size_t lanes, size_t sstride, int layout, size_t *done) {
    *done = 0;
    if (ctx->policy == 1 || layout != IDSP_LANE_MAJOR) return IDSP_HBF_FAST_NOT_APPLICABLE;
    const size_t TO = (size_t)hf::TT >> k;
    const size_t ntiles = n_out / TO;
    const bool ok = ntiles >= 1 && (((uintptr_t)x) & 15) == 0 && ((n_out << k) % 4) == 0;
    if (!ok) return IDSP_HBF_FAST_NOT_APPLICABLE;
    int r;
    switch (k) {
        case 1: r = hf::launch<1>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
        case 2: r = hf::launch<2>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
        case 3: r = hf::launch<3>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
        case 4: r = hf::launch<4>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
        default: r = hf::launch<5>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break;
    }
    if (r == IDSP_OK) *done = ntiles * TO;
    return r;
}

}  // namespace idsp
