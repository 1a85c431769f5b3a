// hbf_fast.cuh -- shared-memory tiled HBF /2^K decimation cascade (f32, lane-major),
// packed f32x2 variant: every stage after the first runs on add/fma.rn.f32x2, two IEEE
// round-to-nearest results per instruction, bit-identical to the scalar operations
// (hbf_fast_scalar.cuh is the scalar variant with the same tiling; both are bit-exact).
//
// A CTA owns NL = 8 lanes for the whole call and walks the time axis in tiles of TT = 512
// input samples per lane (4 CTAs of 128 threads per SM):
//
//   HBM --TMA 1-D bulk copy per lane row (history + 2 KB), mbarrier complete_tx--> raw ring
//   stage 0 : scalar FP32 on the interleaved raw stream.  A work item evaluates the same
//             output positions in the FIRST and in the SECOND half of the tile and writes
//             the two results side by side: element e of a row of stage 1 is the pair
//             (first-half sample e, second-half sample e)
//   stage s : packed arithmetic on those pairs: one instruction advances both halves of
//             the tile; reads E_s / O_s, writes E_{s+1} / O_{s+1} (shared memory)
//   stage K-1 : unpacks and writes the decimated output straight to HBM
//
// Pairing two time halves of the SAME lane (instead of two lanes) keeps 8 rows per CTA, so a
// quarter-warp still touches 8 different rows whose pitch is 4*odd floats: every LDS.128 /
// STS.128 is bank-conflict free, and shared memory per CTA stays at 4 CTAs per SM.
//
// Row layout (elements = float2): lane .x of a row is the stream segment
// [T0 - H, T0 + n/2), lane .y the segment [T0 + n/2 - H, T0 + n) (T0 = tile start, H = filter
// history, n = new samples per tile).  The two segments overlap, so a producer writes the
// last min(H, n/2) samples of the first half twice (.x main slot, .y history slot); the
// carry between tiles is x'[i] = y[i + n/2] and, where H > n/2, y'[i] = y[i + n]
// (the reference's copy_within, src/hbf.rs:183-184, on both segments at once).
//
// A thread loads its window into registers with static indices and evaluates R outputs in
// exactly the reference's order:
//   acc = ((w[2M-1]+w[0])*c0) + ((w[2M-2]+w[1])*c1) + ...  then  + even sample
// (src/hbf.rs:46-68, :178-181), each op individually rounded -> bit-exact.  The ABI state
// (even/odd history per stage) is scattered into the row heads at entry and gathered back
// at exit, so calls can be chained like block() calls on the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "hbf_stages.cuh"
#include "tma_kernels.cuh"

#define IDSP_HBF_FAST_NOT_APPLICABLE 12346

namespace idsp {
namespace hf {

constexpr int NL = 8;  // lanes per CTA
#ifndef HF_NT
#define HF_NT 128
#endif
constexpr int NT = HF_NT;  // threads per CTA
#ifndef HF_TT
#define HF_TT 512
#endif
constexpr int TT = HF_TT;  // raw input samples per lane per tile
constexpr int S = 2;       // raw ring depth
#ifndef HF_MINB
#define HF_MINB 4
#endif
#ifndef HF_RLAST  // pair-outputs per item of a 23-tap stage
#define HF_RLAST 4
#endif
#ifndef HF_ITEMS  // target number of items per packed stage (sets R)
#define HF_ITEMS 64
#endif

__host__ __device__ constexpr int up2(int v) { return (v + 1) & ~1; }
__host__ __device__ constexpr int up4(int v) { return (v + 3) & ~3; }
// pitch in floats: multiple of 4 with pitch/4 odd (conflict-free 16-byte accesses over 8 rows)
__host__ __device__ constexpr int oddpitch(int v) { return (up4(v) / 4) % 2 ? up4(v) : up4(v) + 4; }
__host__ __device__ constexpr int st_m(int K, int s) { return hbf_m(K - 1 - s); }
__host__ __device__ constexpr int st_n(int s) { return TT >> (s + 1); }  // outputs per lane per tile
__host__ __device__ constexpr int st_np(int s) { return st_n(s) / 2; }   // pair-outputs per lane per tile
// outputs (stage 0: per half) / pair-outputs per work item
__host__ __device__ constexpr int st_r(int K, int s) {
    if (s == 0) return 8;
    if (st_m(K, s) > 12) return HF_RLAST;
    int r = NL * st_np(s) / HF_ITEMS;
    return r > 8 ? 8 : (r < 2 ? 2 : r);
}
__host__ __device__ constexpr int raw_h(int K) { return up4(4 * st_m(K, 0) - 2); }
__host__ __device__ constexpr int raw_pitch(int K) { return oddpitch(raw_h(K) + TT); }
// packed rows (s >= 1): history in ELEMENTS (one element = 2 floats = both tile halves)
__host__ __device__ constexpr int he(int K, int s) { return up2(st_m(K, s) - 1); }
__host__ __device__ constexpr int ho(int K, int s) { return up2(2 * st_m(K, s) - 1); }
__host__ __device__ constexpr int pe(int K, int s) { return oddpitch(2 * (he(K, s) + st_np(s))); }  // floats
__host__ __device__ constexpr int po(int K, int s) { return oddpitch(2 * (ho(K, s) + st_np(s))); }
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

// 128-bit shared loads as explicit *volatile* PTX: ptxas narrows a plain `ld.shared.v4`
// whose components are partly unused (windows over the interleaved stream) to scalar
// LDS / LDS.64, and those are 4-8 way bank conflicted over rows of pitch 4*odd floats.
__device__ __forceinline__ float4 lds128(const float *p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(smem_u32(p)));
    return v;
}
typedef unsigned long long f2;  // two packed f32: lo = first tile half, hi = second tile half
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

// One half of a raw (interleaved) stage item: outputs p0 .. p0+R-1 of lane row `row`
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

// One item of a packed stage: erow = [HE hist | np new] elements, orow = [HO hist | np new],
// element = (first half, second half).  p0 (first pair-output, even) is also the element
// offset of the window inside the rows.  Long filters are evaluated in chunks of TC taps,
// each chunk loading only the window pieces it is the first to need, so that the 23-tap
// stage never holds its whole 2 x 48-register window at once.
template <int TI, int R> struct PackedItem {
    static constexpr int M = HbfTaps<TI>::M;
    static constexpr int LEN = 2 * M - 1;
    static constexpr int HE = up2(M - 1), HO = up2(LEN);
    static constexpr int RE = HE - (M - 1), RO = HO - LEN;
    static constexpr int WO = up2(RO + R + 2 * M - 1), WE = up2(RE + R);
    static constexpr int TC = M > 12 ? 12 : M;
    static constexpr int NC = (M + TC - 1) / TC;
    // does chunk c (taps [c*TC, min(M,(c+1)*TC))) read window piece j (elements 2j, 2j+1)?
    __host__ __device__ static constexpr bool needs(int c, int j) {
        const int i0 = c * TC, i1 = (c + 1) * TC < M ? (c + 1) * TC : M;
        const int lo0 = RO + i0, lo1 = RO + i1 - 1 + R - 1;                          // ascending operand
        const int hi0 = RO + 2 * M - 1 - (i1 - 1), hi1 = RO + 2 * M - 1 - i0 + R - 1;  // descending operand
        const int e0 = 2 * j, e1 = 2 * j + 1;
        return (e1 >= lo0 && e0 <= lo1) || (e1 >= hi0 && e0 <= hi1);
    }
    __host__ __device__ static constexpr int first_chunk(int j) {
        for (int c = 0; c < NC; c++)
            if (needs(c, j)) return c;
        return NC;
    }
    __device__ __forceinline__ static void run(const float *erow, const float *orow, int p0, f2 nz,
                                               f2 (&y)[R]) {
        f2 wo[WO], we[WE], acc[R];
#pragma unroll
        for (int c = 0; c < NC; c++) {
#pragma unroll
            for (int j = 0; j < WO / 2; j++)
                if (first_chunk(j) == c) lds_2f2(orow + 2 * p0 + 4 * j, wo[2 * j], wo[2 * j + 1]);
#pragma unroll
            for (int q = 0; q < R; q++) {
#pragma unroll
                for (int i = c * TC; i < (c + 1) * TC && i < M; i++) {
                    const float ci = HbfTaps<TI>::c(i);
                    const f2 t = mul2(add2(wo[RO + q + 2 * M - 1 - i], wo[RO + q + i]), pk(ci, ci), nz);
                    acc[q] = i == 0 ? t : add2(acc[q], t);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < WE / 2; j++) lds_2f2(erow + 2 * p0 + 4 * j, we[2 * j], we[2 * j + 1]);
#pragma unroll
        for (int q = 0; q < R; q++) y[q] = add2(acc[q], we[RE + q]);
    }
};

// Scatter R consecutive pair-outputs (p0 multiple of R, R even) of stage s-1 into E_s / O_s.
// e_row / o_row point at the row heads; H_E / H_O = history elements, NP = new elements per
// tile.  Output p goes to element p/2 of the even (p even) or odd (p odd) row; first-half
// results that are history of the second half are written a second time into lane .y.
template <int R, int HE_, int HO_, int NP>
__device__ __forceinline__ void put_packed(float *e_row, float *o_row, int p0, const f2 (&y)[R]) {
    float *e_new = e_row + 2 * HE_ + p0, *o_new = o_row + 2 * HO_ + p0;  // element (p0/2) -> float offset p0
    if constexpr (R == 8) {
        sts_2f2(e_new, y[0], y[2]);
        sts_2f2(e_new + 4, y[4], y[6]);
        sts_2f2(o_new, y[1], y[3]);
        sts_2f2(o_new + 4, y[5], y[7]);
    } else if constexpr (R == 4) {
        sts_2f2(e_new, y[0], y[2]);
        sts_2f2(o_new, y[1], y[3]);
    } else {
        static_assert(R == 2, "R must be 2, 4 or 8");
        sts_f2(e_new, y[0]);
        sts_f2(o_new, y[1]);
    }
    // u = element index inside the first half; it is also .y history element u + H - NP when >= 0
    const int u0 = p0 / 2;
    if (u0 + R / 2 > NP - HE_) {
#pragma unroll
        for (int q = 0; q < R; q += 2) {
            const int i = u0 + q / 2 + HE_ - NP;
            float a, b;
            unpk(y[q], a, b);
            if (i >= 0) e_row[2 * i + 1] = a;
        }
    }
    if (u0 + R / 2 > NP - HO_) {
#pragma unroll
        for (int q = 1; q < R; q += 2) {
            const int i = u0 + q / 2 + HO_ - NP;
            float a, b;
            unpk(y[q], a, b);
            if (i >= 0) o_row[2 * i + 1] = a;
        }
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

// Carry one kind of row (E or O) of stage s for all NL lanes: x'[i] = y[i + NP] for i < H,
// y'[i] = y[i + 2 NP] for i < H - NP.  A whole row is handled inside one warp (one 16-byte
// piece = 2 elements per thread), everything is read before anything is written.
template <int H, int NP, int PITCH>
__device__ __forceinline__ void carry_kind(float *rows, int wsel, int nw, int lid) {
    constexpr int C = H / 2;  // pieces per row head
    static_assert(C <= 32 && H % 2 == 0 && NP % 2 == 0, "carry layout");
    constexpr int LP = C > 16 ? 1 : C > 8 ? 2 : C > 4 ? 4 : 8;  // lanes per warp pass
    constexpr int CP = 32 / LP;
    const int sub = lid / CP, j = lid % CP;
    for (int pass = wsel; pass * LP < NL; pass += nw) {
        const int lane = pass * LP + sub;
        const bool act = j < C && lane < NL;
        float *head = rows + lane * PITCH + 4 * j;
        float4 p1 = make_float4(0.f, 0.f, 0.f, 0.f), p2 = p1;
        if (act) {
            p1 = lds128(head + 2 * NP);
            if (H > NP && 2 * j < H - NP) p2 = lds128(head + 4 * NP);
        }
        __syncwarp();
        if (act) *reinterpret_cast<float4 *>(head) = make_float4(p1.y, p2.y, p1.w, p2.w);
        __syncwarp();
    }
}
template <int K, int s>
__device__ __forceinline__ void carry_rows(float *sm, int wsel, int nw, int lid) {
    carry_kind<he(K, s), st_np(s), pe(K, s)>(sm + off_e(K, s), wsel, nw, lid);
    carry_kind<ho(K, s), st_np(s), po(K, s)>(sm + off_o(K, s), wsel, nw, lid);
}

template <int K, int s> struct StageRun {
    static constexpr int TI = K - 1 - s;
    static constexpr int R = st_r(K, s);
    static constexpr int NP = st_np(s);
    static constexpr int ITEMS = NL * NP / R;
    // Warps that hold items of this stage; odd stages take the upper warps so that, over the
    // CTAs of an SM, every scheduler gets work.  The other warps carry rows s-1 meanwhile.
    static constexpr int NW = NT / 32;
    static constexpr int NWA = ITEMS >= NT ? NW : (ITEMS + 31) / 32;
    static constexpr int W0 = (s & 1) ? NW - NWA : 0;

    // runs packed stage s (1 <= s <= K-1) for one tile
    __device__ __forceinline__ static void run(float *sm, int tid, int nl, float *y, size_t ystride,
                                               size_t yoff, size_t lane0, f2 nz) {
        const int vt = tid - 32 * W0;
        if (vt >= 0 && vt < 32 * NWA) {
            const float *E = sm + off_e(K, s);
            const float *O = sm + off_o(K, s);
            for (int idx = vt; idx < ITEMS; idx += 32 * NWA) {
                const int lane = idx % NL, p0 = (idx / NL) * R;
                f2 out[R];
                PackedItem<TI, R>::run(E + lane * pe(K, s), O + lane * po(K, s), p0, nz, out);
                if constexpr (s == K - 1) {
                    float a[R], b[R];
#pragma unroll
                    for (int q = 0; q < R; q++) unpk(out[q], a[q], b[q]);
                    if (lane < nl) {
                        float *dst = y + (lane0 + lane) * ystride + yoff + p0;
                        store_out<R>(dst, a);
                        store_out<R>(dst + NP, b);
                    }
                } else {
                    put_packed<R, he(K, s + 1), ho(K, s + 1), st_np(s + 1)>(
                        sm + off_e(K, s + 1) + lane * pe(K, s + 1), sm + off_o(K, s + 1) + lane * po(K, s + 1), p0, out);
                }
            }
        }
        if constexpr (s >= 2) {  // rows s-1 were consumed in the previous phase
            const int warp = tid >> 5;
            if constexpr (NWA < NW) {
                if (vt < 0 || vt >= 32 * NWA) carry_rows<K, s - 1>(sm, vt < 0 ? warp : warp - NWA, NW - NWA, tid & 31);
            } else {
                carry_rows<K, s - 1>(sm, warp, NW, tid & 31);
            }
        }
    }
};

// ABI state <-> shared-memory histories (see header comment of include/idsp_b200.h).
// Raw history lives at row[roff .. roff+HR): roff = 0 (head) on entry, TT (tail of the last
// tile) on exit.  Packed rows: word w of a history of Mh values is lane .x of head element
// H - Mh + w and, on entry, also lane .y of element H - Mh + w - NP when that is >= 0.
template <int K, int s, bool LOAD> struct StateIO {
    __device__ __forceinline__ static void run(float *sm, float *st, size_t sstride, size_t lane0, int nl,
                                               int tid, int rawbuf, int roff) {
        if constexpr (s < K) {
            constexpr int M = st_m(K, s);
            constexpr int LEN = 2 * M - 1;
            constexpr int WORDS = 3 * M - 2;
            float *stw = st + (size_t)st_word(K, s) * sstride + lane0;
            for (int idx = tid; idx < WORDS * NL; idx += NT) {
                const int lane = idx % NL, w = idx / NL;
                if (lane >= nl) continue;
                if constexpr (s == 0) {
                    constexpr int HR = raw_h(K);
                    float *row = sm + (rawbuf * NL + lane) * raw_pitch(K) + roff;
                    float *p = w < M - 1 ? row + (HR - 2 * M + 2 + 2 * w) : row + (HR - 4 * M + 3 + 2 * (w - (M - 1)));
                    if constexpr (LOAD) *p = stw[(size_t)w * sstride + lane];
                    else stw[(size_t)w * sstride + lane] = *p;
                } else {
                    constexpr int NP = st_np(s);
                    float *row = w < M - 1 ? sm + off_e(K, s) + lane * pe(K, s) : sm + off_o(K, s) + lane * po(K, s);
                    const int e = w < M - 1 ? he(K, s) - (M - 1) + w : ho(K, s) - LEN + (w - (M - 1));
                    if constexpr (LOAD) {
                        const float v = stw[(size_t)w * sstride + lane];
                        row[2 * e] = v;
                        if (e - NP >= 0) row[2 * (e - NP) + 1] = v;
                    } else {
                        stw[(size_t)w * sstride + lane] = row[2 * e];
                    }
                }
            }
            StateIO<K, s + 1, LOAD>::run(sm, st, sstride, lane0, nl, tid, rawbuf, roff);
        }
    }
};

template <int K>
__global__ void __launch_bounds__(NT, HF_MINB)
hbf_dec_fast_kernel(float *st, const float *x, float *y, size_t n_out, size_t ntiles, size_t lanes,
                    size_t sstride, f2 nz /* (-0.0f, -0.0f), see mul2() */) {
    constexpr int TI0 = K - 1;
    constexpr int R0 = st_r(K, 0);
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
    StateIO<K, 0, true>::run(sm, st, sstride, lane0, nl, tid, 0, 0);
    // generic-proxy writes above (zero fill) precede async-proxy (TMA) writes to the same rows
    fence_async_smem();
    __syncthreads();

    // Tile t >= 1 is fetched together with the HR samples in front of it (they are in L2 from
    // the previous tile), so the raw history never has to be copied between ring buffers;
    // tile 0 takes its history from the ABI state (scattered into buffer 0 above).
    auto issue = [&](size_t tile) {  // executed by warp 0
        const int b = (int)(tile % S);
        const uint32_t bar = smem_u32(&bars[b]);
        const uint32_t hist = tile ? HR : 0;
        if ((tid & 31) == 0) mbar_expect_tx(bar, (uint32_t)(nl * (TT + hist) * 4));
        __syncwarp();
        if (tid < nl)
            bulk_load_1d(smem_u32(sm + (b * NL + tid) * PR + HR - hist),
                         x + (lane0 + tid) * n_in + tile * TT - hist, (TT + hist) * 4, bar);
    };
    if (tid < 32) {
#pragma unroll
        for (int b = 0; b < S; b++)
            if ((size_t)b < ntiles) issue(b);
    }

    for (size_t i = 0; i < ntiles; i++) {
        const int b = (int)(i % S);
        mbar_wait(smem_u32(&bars[b]), (uint32_t)((i / S) & 1));
        // ---- phase 0: raw interleaved stream (scalar FP32)
        {
            const float *raw = sm + b * NL * PR;
            if constexpr (K == 1) {  // single stage: straight to HBM
                constexpr int ITEMS = NL * st_n(0) / R0;
                for (int idx = tid; idx < ITEMS; idx += NT) {
                    const int lane = idx % NL, p0 = (idx / NL) * R0;
                    float out[R0];
                    RawItem<TI0, R0>::run(raw + lane * PR, p0, out);
                    if (lane < nl) store_out<R0>(y + (lane0 + lane) * n_out + i * TO + p0, out);
                }
            } else {  // same positions of both tile halves -> packed E_1 / O_1
                constexpr int ITEMS = NL * st_np(0) / R0;
                for (int idx = tid; idx < ITEMS; idx += NT) {
                    const int lane = idx % NL, p0 = (idx / NL) * R0;
                    float a[R0], bb[R0];
                    RawItem<TI0, R0>::run(raw + lane * PR, p0, a);
                    RawItem<TI0, R0>::run(raw + lane * PR + TT / 2, p0, bb);
                    f2 out[R0];
#pragma unroll
                    for (int q = 0; q < R0; q++) out[q] = pk(a[q], bb[q]);
                    put_packed<R0, he(K, 1), ho(K, 1), st_np(1)>(sm + off_e(K, 1) + lane * pe(K, 1),
                                                                 sm + off_o(K, 1) + lane * po(K, 1), p0, out);
                }
                // rows K-1 of the previous tile (last read in its final phase, next written in phase K-2 >= 1)
                if constexpr (K >= 3) {
                    if (i > 0) carry_rows<K, K - 1>(sm, tid >> 5, NT / 32, tid & 31);
                }
            }
        }
        __syncthreads();
        // ---- raw buffer b is free again: refill it
        if (tid < 32 && i + S < ntiles) issue(i + S);
        // ---- packed phases 1 .. K-1 (phase s also carries rows s-1)
        if constexpr (K >= 2) { StageRun<K, 1>::run(sm, tid, nl, y, n_out, i * TO, lane0, nz); __syncthreads(); }
        if constexpr (K >= 3) { StageRun<K, 2>::run(sm, tid, nl, y, n_out, i * TO, lane0, nz); __syncthreads(); }
        if constexpr (K >= 4) { StageRun<K, 3>::run(sm, tid, nl, y, n_out, i * TO, lane0, nz); __syncthreads(); }
        if constexpr (K >= 5) { StageRun<K, 4>::run(sm, tid, nl, y, n_out, i * TO, lane0, nz); __syncthreads(); }
        if constexpr (K == 2) {  // rows 1 are written again in the very next phase: carry them now
            carry_rows<K, 1>(sm, tid >> 5, NT / 32, tid & 31);
            __syncthreads();
        }
    }
    if constexpr (K >= 3) {
        carry_rows<K, K - 1>(sm, tid >> 5, NT / 32, tid & 31);
        __syncthreads();
    }
    // the raw history of the stream is the tail of the last tile's buffer
    StateIO<K, 0, false>::run(sm, st, sstride, lane0, nl, tid, (int)((ntiles - 1) % S), TT);
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
