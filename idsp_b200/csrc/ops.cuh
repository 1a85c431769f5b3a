// ops.cuh -- per-lane sample operators ("Op"s) executed by the lane kernels.
//
// One Op instance lives in the registers of the thread that owns a filter lane:
//   load()  reads the lane's state words (SoA: st[word*stride + lane]),
//   step()  consumes one input sample and returns one output sample,
//   store() writes the state back.
// The arithmetic restates the reference exactly (wrapping integer accumulators,
// arithmetic shifts, per-op rounded floats, no FMA: the library is compiled with
// -fmad=false).  Reference file:line is cited at each Op.
#pragma once
#include <type_traits>
#include <stdint.h>

#include "tables.cuh"

namespace idsp {

// --------------------------------------------------------------------------
// lookup tables (build.rs:9-69) in global memory, read through L1 (__ldg)
// --------------------------------------------------------------------------
__device__ const uint32_t g_cossin_lut[128] = IDSP_COSSIN_TABLE_INIT;
__device__ const uint2 g_divi_tab[16] = IDSP_ATAN2_DIVI_PAIR_INIT;  // (base, slope bits), build.rs:46-69

template <class T> struct Wide;
template <> struct Wide<int8_t> { using A = int16_t; using UA = uint16_t; using UT = uint8_t; };
template <> struct Wide<int16_t> { using A = int32_t; using UA = uint32_t; using UT = uint16_t; };
template <> struct Wide<int32_t> { using A = int64_t; using UA = uint64_t; using UT = uint32_t; };
template <> struct Wide<int64_t> { using A = __int128; using UA = unsigned __int128; using UT = uint64_t; };

template <class T> struct is_float { static constexpr bool value = false; };
template <> struct is_float<float> { static constexpr bool value = true; };
template <> struct is_float<double> { static constexpr bool value = true; };

// Defaults of the optional Op hooks used by the TMA kernels (tma_kernels.cuh)
struct OpHooks {
    static constexpr bool TUNABLE = false;
    static constexpr bool HEAVY = false;  // compute-bound: lane-major TMA kernels trade tile size for resident warps
    static constexpr bool LM_SMALL = false;  // in between: fewer lane-major stages of the same tile
    static constexpr int SMEM_EXTRA_WORDS = 0;
    template <class P> __device__ __forceinline__ static void init_smem(const P &, uint32_t *, int, int) {}
    template <class P> __device__ __forceinline__ void bind(const P &, const uint32_t *) {}
};

// a * b + c with a, b i32 and c i64: one IMAD.WIDE (the compiler turns an i64 product of a
// sign-extended register and a kernel parameter into a 5-instruction 64x32 multiply)
__device__ __forceinline__ int64_t mad_wide(int32_t a, int32_t b, int64_t c) {
    int64_t r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
// (a * b) >> 32 whose result ptxas cannot fold into a following add / subtract: it turns mul.hi + add into
// IMAD.HI with a 64-bit addend {0, c}, which costs two register moves per fold to build the pair -- more work
// on the multiplier pipe than the add it saves (the lock-in kernels ran 12 such moves per sample).
__device__ __forceinline__ int32_t mulhi_opaque(int32_t a, int32_t b) {
    int32_t lo, hi;
    asm volatile("{ .reg .b64 t; mul.wide.s32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
    (void)lo;
    return hi;
}
// num_traits::clamp (src/iir/biquad.rs:400): `<`/`>` compares so NaN passes through
template <class T> __device__ __forceinline__ T clamp_nt(T v, T lo, T hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

// --------------------------------------------------------------------------
// y0 = (b0 x0 + b1 x1 + b2 x2 + a1 y1 + a2 y2) >> F      src/iir/biquad.rs:366-383
// Q*T -> wide accu: dsp-fixedpoint/src/ops.rs:91-97, lib.rs:310-312
// accu.as_() = (acc >> F) as T: num_traits_impl.rs:74-104, lib.rs:297-299
// --------------------------------------------------------------------------
// 32-bit multiply-accumulate whose ORDER the compiler keeps (plain C++ integer sums are re-associated).
__device__ __forceinline__ int32_t madlo(int32_t a, int32_t b, int32_t c) {
    int32_t d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <class T, bool ISF = is_float<T>::value> struct Sos;

template <class T> struct Sos<T, false> {
    using A = typename Wide<T>::A;
    using UA = typename Wide<T>::UA;
    using UT = typename Wide<T>::UT;
    // The five products are summed in wrapping arithmetic (wrapA), which is associative: the
    // order is free.  8- / 16-bit samples: one 32-bit multiply-add chain, the operands that exist earliest
    // first and a1*y1 (the recurrence) last; the accumulator wraps at 16 / 32 bits = the low bits of the
    // 32-bit sum (i8 lane-major 1 411 -> 1 814 GSa/s, i16 frame-major 870 -> 934).  i32: the compiler's
    // single IMAD.WIDE accumulate chain is kept -- forcing an order makes ptxas split every multiply-add into
    // IMAD.WIDE + two carry adds (DF1 i32 805 -> 763 GSa/s, Cascade<4> unchanged: that kernel is bound by
    // the issue rate of IMAD.WIDE, not by the length of the chain).
    __device__ __forceinline__ static T eval(const T *ba, int F, T x0, T x1, T x2, T y1, T y2) {
        UA acc;
        if constexpr (sizeof(T) < 4) {
            const int32_t a32 = madlo(ba[3], y1, madlo(ba[0], x0, madlo(ba[1], x1, madlo(ba[4], y2, madlo(ba[2], x2, 0)))));
            acc = (UA)(uint32_t)a32;
        } else {
            acc = (UA)((A)ba[0] * (A)x0) + (UA)((A)ba[1] * (A)x1) + (UA)((A)ba[2] * (A)x2) +
                  (UA)((A)ba[3] * (A)y1) + (UA)((A)ba[4] * (A)y2);
        }
        A q = F >= 0 ? (A)((A)acc >> F) : (A)(UA)(acc << (-F));
        return (T)(UT)(UA)q;
    }
    __device__ __forceinline__ static T add(T a, T b) { return (T)(UT)((UT)a + (UT)b); }
};
// i32 with 0 <= F < 32: low word of (acc >> F) is one funnel shift
struct SosI32Fast {
    __device__ __forceinline__ static int32_t eval(const int32_t *ba, int F, int32_t x0,
                                                   int32_t x1, int32_t x2, int32_t y1,
                                                   int32_t y2) {
        uint64_t acc = (uint64_t)((int64_t)ba[0] * x0) + (uint64_t)((int64_t)ba[1] * x1) +
                       (uint64_t)((int64_t)ba[2] * x2) + (uint64_t)((int64_t)ba[3] * y1) +
                       (uint64_t)((int64_t)ba[4] * y2);
        return (int32_t)__funnelshift_r((uint32_t)acc, (uint32_t)(acc >> 32), F);
    }
};
template <class T> struct Sos<T, true> {
    __device__ __forceinline__ static T eval(const T *ba, int, T x0, T x1, T x2, T y1, T y2) {
        // left-to-right, each op rounded (no FMA contraction: -fmad=false)
        return ba[0] * x0 + ba[1] * x1 + ba[2] * x2 + ba[3] * y1 + ba[4] * y2;
    }
    __device__ __forceinline__ static T add(T a, T b) { return a + b; }
};

// DF1 + optional clamp (src/iir/biquad.rs:366-404). MODE: 0 generic F, 1 i32 0<=F<32
template <class T, bool CLAMP, int MODE = 0> struct Df1Op {
    using In = T;
    using Out = T;
    static constexpr bool TUNABLE = true;  // tuning builds sweep tile shapes for this Op
    static constexpr bool HEAVY = false;
    static constexpr bool LM_SMALL = false;
    static constexpr int SMEM_EXTRA_WORDS = 0;
    struct Params {
        T ba[5];
        int F;
        T u, mn, mx;
        T *st;
    };
    __device__ __forceinline__ static void init_smem(const Params &, uint32_t *, int, int) {}
    __device__ __forceinline__ void bind(const Params &, const uint32_t *) {}
    T x1, x2, y1, y2;
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        x1 = p.st[lane];
        x2 = p.st[stride + lane];
        y1 = p.st[2 * stride + lane];
        y2 = p.st[3 * stride + lane];
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.st[lane] = x1;
        p.st[stride + lane] = x2;
        p.st[2 * stride + lane] = y1;
        p.st[3 * stride + lane] = y2;
    }
    __device__ __forceinline__ T step(const Params &p, T x0) {
        T y0;
        if constexpr (MODE == 1)
            y0 = SosI32Fast::eval(p.ba, p.F, x0, x1, x2, y1, y2);
        else
            y0 = Sos<T>::eval(p.ba, p.F, x0, x1, x2, y1, y2);
        x2 = x1;
        x1 = x0;
        y2 = y1;
        if constexpr (CLAMP) y0 = clamp_nt<T>(Sos<T>::add(y0, p.u), p.mn, p.mx);  // biquad.rs:399-402
        y1 = y0;
        return y0;
    }
};

// P = 4 / sizeof(T) adjacent lanes of a 1- or 2-byte Op packed into one 32-bit word: lets the
// frame-major TMA kernels (4-byte samples) carry i8 / i16 lanes, one word = P lanes of one frame,
// and gives every thread P independent recurrences.  `lane` counts words, state stays SoA per lane.
template <class Op> struct PackedOp : OpHooks {
    using T = typename Op::In;
    using UT = typename Wide<T>::UT;
    static_assert(sizeof(T) < 4 && sizeof(typename Op::Out) == sizeof(T), "1- or 2-byte samples");
    static constexpr int P = 4 / (int)sizeof(T);
    using In = int32_t;
    using Out = int32_t;
    using Params = typename Op::Params;
    Op op[P];
    __device__ __forceinline__ void load(const Params &p, size_t word, size_t stride) {
#pragma unroll
        for (int k = 0; k < P; k++) op[k].load(p, word * P + k, stride);
    }
    __device__ __forceinline__ void store(const Params &p, size_t word, size_t stride) const {
#pragma unroll
        for (int k = 0; k < P; k++) op[k].store(p, word * P + k, stride);
    }
    __device__ __forceinline__ int32_t step(const Params &p, int32_t w) {
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < P; k++) {
            const T o = op[k].step(p, (T)(UT)((uint32_t)w >> (8 * sizeof(T) * k)));
            r |= (uint32_t)(UT)o << (8 * sizeof(T) * k);
        }
        return (int32_t)r;
    }
};

// Cascade<[Biquad;N]> on DirectForm<T,N> (src/iir/biquad.rs:339-364)
template <class T, int N, int MODE = 0> struct CascadeOp : OpHooks {
    // exactly N sections (the entry point dispatches on nsec): every index below is static, so
    // the 2 + 2N delay values stay in registers.  MODE 1 (i32, 0 <= F < 32, chosen by the entry point):
    // funnel-shift quantiser with no per-section test of F -- a branch per section, even a uniform one,
    // keeps the scheduler from overlapping the sections of consecutive samples.
    using In = T;
    using Out = T;
    static_assert(MODE == 0 || std::is_same<T, int32_t>::value, "MODE 1 is the i32 fast path");
    static constexpr bool HEAVY = N >= 2;
    struct Params {
        T ba[N][5];
        int F;
        int nsec;
        T *st;
    };
    T d[2 + 2 * N];  // [x0,x1,y[0][0],y[0][1],...]
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
#pragma unroll
        for (int w = 0; w < 2 + 2 * N; w++) d[w] = p.st[(size_t)w * stride + lane];
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
#pragma unroll
        for (int w = 0; w < 2 + 2 * N; w++) p.st[(size_t)w * stride + lane] = d[w];
    }
    template <int S> __device__ __forceinline__ T section(const Params &p, T x0) {
        T y0;
        if constexpr (MODE == 1)
            y0 = SosI32Fast::eval(p.ba[S], p.F, x0, d[2 * S], d[2 * S + 1], d[2 * S + 2], d[2 * S + 3]);
        else
            y0 = Sos<T>::eval(p.ba[S], p.F, x0, d[2 * S], d[2 * S + 1], d[2 * S + 2], d[2 * S + 3]);
        d[2 * S + 1] = d[2 * S];
        d[2 * S] = x0;
        if constexpr (S + 1 < N) return section<S + 1>(p, y0);
        else return y0;
    }
    __device__ __forceinline__ T step(const Params &p, T x0) {
        x0 = section<0>(p, x0);
        d[2 * N + 1] = d[2 * N];
        d[2 * N] = x0;
        return x0;
    }
};

// DF2T (src/iir/biquad.rs:418-440)
template <class T, bool CLAMP> struct Df2tOp : OpHooks {
    using In = T;
    using Out = T;
    struct Params {
        T ba[5];
        T u, mn, mx;
        T *st;
    };
    T s0, s1;
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        s0 = p.st[lane];
        s1 = p.st[stride + lane];
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.st[lane] = s0;
        p.st[stride + lane] = s1;
    }
    __device__ __forceinline__ T step(const Params &p, T x0) {
        T y0 = s0 + p.ba[0] * x0;
        if constexpr (CLAMP) y0 = clamp_nt<T>(y0 + p.u, p.mn, p.mx);
        s0 = s1 + p.ba[1] * x0 + p.ba[3] * y0;
        s1 = p.ba[2] * x0 + p.ba[4] * y0;
        return y0;
    }
};

// DirectForm1Wide (src/iir/biquad.rs:445-480)
template <bool CLAMP> struct Df1WideOp : OpHooks {
    using In = int32_t;
    using Out = int32_t;
    static constexpr bool HEAVY = true;  // 7 wide MACs per sample: per-warp pipelines measure 516 -> 561 GSa/s frame-major
    struct Params {
        int32_t ba[5];
        int F;
        int32_t u, mn, mx;
        int32_t *st;
    };
    int32_t x1, x2;
    int64_t y1, y2;
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        x1 = p.st[lane];
        x2 = p.st[stride + lane];
        y1 = (int64_t)(((uint64_t)(uint32_t)p.st[3 * stride + lane] << 32) |
                       (uint32_t)p.st[2 * stride + lane]);
        y2 = (int64_t)(((uint64_t)(uint32_t)p.st[5 * stride + lane] << 32) |
                       (uint32_t)p.st[4 * stride + lane]);
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.st[lane] = x1;
        p.st[stride + lane] = x2;
        p.st[2 * stride + lane] = (int32_t)(uint32_t)y1;
        p.st[3 * stride + lane] = (int32_t)(y1 >> 32);
        p.st[4 * stride + lane] = (int32_t)(uint32_t)y2;
        p.st[5 * stride + lane] = (int32_t)(y2 >> 32);
    }
    __device__ __forceinline__ int32_t step(const Params &p, int32_t x0) {
        uint64_t acc = (uint64_t)((int64_t)p.ba[0] * x0) + (uint64_t)((int64_t)p.ba[1] * x1) +
                       (uint64_t)((int64_t)p.ba[2] * x2);
        x2 = x1;
        x1 = x0;
        acc += (uint64_t)(((int64_t)(uint64_t)(uint32_t)y1 * (int64_t)p.ba[3]) >> 32);
        acc += (uint64_t)((int64_t)(int32_t)(y1 >> 32) * (int64_t)p.ba[3]);
        acc += (uint64_t)(((int64_t)(uint64_t)(uint32_t)y2 * (int64_t)p.ba[4]) >> 32);
        acc += (uint64_t)((int64_t)(int32_t)(y2 >> 32) * (int64_t)p.ba[4]);
        acc <<= (32 - p.F);
        y2 = y1;
        y1 = (int64_t)acc;
        int32_t y0 = (int32_t)((int64_t)acc >> 32);
        if constexpr (CLAMP) {  // biquad.rs:474-480
            y0 = clamp_nt<int32_t>((int32_t)((uint32_t)y0 + (uint32_t)p.u), p.mn, p.mx);
            y1 = (int64_t)(((uint64_t)(int64_t)y0 << 32) | (uint64_t)(uint32_t)y1);
        }
        return y0;
    }
};

// DirectForm1Dither (src/iir/biquad.rs:484-538)
template <bool CLAMP> struct Df1DitherOp : OpHooks {
    using In = int32_t;
    using Out = int32_t;
    static constexpr bool LM_SMALL = true;
    struct Params {
        int32_t ba[5];
        int F;
        int32_t u, mn, mx;
        int32_t *st;
    };
    int32_t x1, x2, y1, y2;
    uint32_t e;
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        x1 = p.st[lane];
        x2 = p.st[stride + lane];
        y1 = p.st[2 * stride + lane];
        y2 = p.st[3 * stride + lane];
        e = (uint32_t)p.st[4 * stride + lane];
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.st[lane] = x1;
        p.st[stride + lane] = x2;
        p.st[2 * stride + lane] = y1;
        p.st[3 * stride + lane] = y2;
        p.st[4 * stride + lane] = (int32_t)e;
    }
    __device__ __forceinline__ int32_t step(const Params &p, int32_t x0) {
        // acc <<= 32 - F; e = (acc as u32) >> (32 - F); y0 = (acc >> 32) as i32   (biquad.rs:520-526)
        // == y0 = bits [F, F+32) of acc (one funnel shift), e = low F bits of acc (0 <= F < 32)
        int64_t acc = mad_wide(p.ba[0], x0, (int64_t)(uint64_t)e);
        acc = mad_wide(p.ba[1], x1, acc);
        acc = mad_wide(p.ba[2], x2, acc);
        acc = mad_wide(p.ba[3], y1, acc);
        acc = mad_wide(p.ba[4], y2, acc);
        e = (uint32_t)acc & ((1u << p.F) - 1u);
        int32_t y0 = (int32_t)__funnelshift_r((uint32_t)acc, (uint32_t)((uint64_t)acc >> 32), (uint32_t)p.F);
        x2 = x1;
        x1 = x0;
        y2 = y1;
        if constexpr (CLAMP) y0 = clamp_nt<int32_t>((int32_t)((uint32_t)y0 + (uint32_t)p.u), p.mn, p.mx);
        y1 = y0;
        return y0;
    }
};

// --------------------------------------------------------------------------
// cossin (src/cossin.rs:14-67). LUT pointer may be global (__ldg) or shared.
// --------------------------------------------------------------------------
template <bool SMEM_LUT>
__device__ __forceinline__ void cossin_dev(const uint32_t *lut, int32_t phase, int32_t &co,
                                           int32_t &so) {
    uint32_t octant = (uint32_t)phase;
    if (octant & (1u << 29)) phase = ~phase;
    phase = (int32_t)((((uint32_t)phase) << 3) >> 10);
    uint32_t lookup = SMEM_LUT ? lut[phase >> 15] : __ldg(lut + (phase >> 15));
    phase &= (1 << 15) - 1;
    phase -= 1 << 14;
    int32_t dphi = (phase * 51471) >> 16;
    int32_t c = (int32_t)(lookup & 0xffffu) + (1 << 16);
    int32_t s = (int32_t)(lookup >> 16);
    int32_t dcos = (s * dphi) >> 7;
    int32_t dsin = (c * dphi) >> 8;
    c = (c << 14) - dcos;
    s = (s << 15) + dsin;
    octant ^= octant >> 1;
    if (octant & (1u << 29)) { int32_t t = c; c = s; s = t; }
    if (octant & (1u << 30)) c = -c;
    if (octant & (1u << 31)) s = -s;
    co = c;
    so = s;
}

// Same function on a pre-expanded table: entry i of the staged shared-memory table holds
// c14 = ((lut & 0xffff) + 65536) << 14 and s15 = (lut >> 16) << 15, i.e. the two values the
// reference forms before the interpolation (src/cossin.rs:44-60).  With d10 = dphi << 10,
//   (s * dphi) >> 7 == (s15 * d10) >> 32   and   (c * dphi) >> 8 == (c14 * d10) >> 32
// exactly (floor of the same rational; |d10| < 2^24, c14 < 2^31), so each correction is one
// IMAD.HI and the unpack / shift instructions disappear.  Bit-exact with cossin_dev
// (tests/test_gpu_nco.py sweeps it against the oracle).
// REP > 1: the table is replicated REP times, entry-major (entry i, copy r at word 2 * (REP * i + r)); a
// thread reads copy `threadIdx.x % REP` only.  With REP = 16 an entry spans all 32 banks and the 16 threads
// of a half-warp (one 8-byte access phase) hit 16 different bank pairs whatever their entries are: the
// data-dependent lookups are conflict free (plain table: 59 % of the load wavefronts were replays).
// Measured (profiles/r2_cossin_table_replication.log): the `cossin` map kernel gains 3.5 % (455 -> 471 GSa/s);
// the lock-in kernels do not (336 -> 326 GSa/s with 16 copies shared by 4 warps, (x, phase) 287 -> 217): their
// limiter is the ALU pipe, not the shared-memory replays, and the table costs residency.  So: 16 copies in the
// map kernel, one in the lock-in kernels.
#ifndef IDSP_COSSIN_REP
#define IDSP_COSSIN_REP 16
#endif
#ifndef IDSP_LOCKIN_LUT_REP
#define IDSP_LOCKIN_LUT_REP 1
#endif
template <int REP = 1>
__device__ __forceinline__ void cossin_expand_lut(const uint32_t *lut, uint32_t *table, int tid, int nthreads) {
    for (int i = tid; i < 128 * REP; i += nthreads) {
        const uint32_t w = lut[i / REP];
        table[2 * i] = ((w & 0xffffu) + 65536u) << 14;
        table[2 * i + 1] = (w >> 16) << 15;
    }
}
// WIDE_HI: the two corrections as unfoldable high-word products (mulhi_opaque): pays in the lock-in kernels
// (multiplier pipe, 365 -> 370 GSa/s), not in the HBM-bound map kernel (468 -> 450)
template <int REP = 1, bool WIDE_HI = false>
__device__ __forceinline__ void cossin_dev_x(const uint32_t *table, int32_t phase, int32_t &co, int32_t &so) {
    uint32_t octant = (uint32_t)phase;
    if (octant & (1u << 29)) phase = ~phase;
    const uint32_t ph = (((uint32_t)phase) << 3) >> 10;
    const uint2 e = *reinterpret_cast<const uint2 *>(table + 2 * REP * (ph >> 15));
    const int32_t frac = (int32_t)(ph & 0x7fffu) - (1 << 14);
    const int32_t d10 = ((frac * 51471) >> 6) & ~0x3ff;
    int32_t c, s;
    if constexpr (WIDE_HI) {
        c = (int32_t)e.x - mulhi_opaque((int32_t)e.y, d10);
        s = (int32_t)e.y + mulhi_opaque((int32_t)e.x, d10);
    } else {
        c = (int32_t)e.x - __mulhi((int32_t)e.y, d10);
        s = (int32_t)e.y + __mulhi((int32_t)e.x, d10);
    }
    octant ^= octant >> 1;
    if (octant & (1u << 29)) { int32_t t = c; c = s; s = t; }
    if (octant & (1u << 30)) c = -c;
    if (octant & (1u << 31)) s = -s;
    co = c;
    so = s;
}

// --------------------------------------------------------------------------
// atan2 (src/atan2.rs:7-82)
// --------------------------------------------------------------------------
// Written branch-free and with explicit 32x32 -> 64 / high-word multiplies so that (a) one call is
// ~50 instead of ~80 SASS instructions (the compiler expands the reference's mixed signed / unsigned
// i64 products into multi-instruction sequences) and (b) the unrolled frame loops of the lane kernels
// can interleave several independent calls (a branch per call serialises them).  Bit-exact with the
// straightforward transcription (tests/test_gpu_nco.py sweeps it against the oracle).
__device__ __forceinline__ uint32_t mul_q31(uint32_t x, uint32_t y) {
    const uint64_t p = (uint64_t)x * (uint64_t)y;  // IMAD.WIDE.U32
    return __funnelshift_r((uint32_t)p, (uint32_t)(p >> 32), 31);
}
// 16 x (base, slope) reciprocal seeds, one 8-byte load per lookup: global (__ldg) or staged in shared memory
template <bool SMEM_TAB>
__device__ __forceinline__ uint32_t divi_dev(const uint2 *tab, uint32_t y, uint32_t x) {
    const uint32_t x_in = x;  // x == 0 (then y == 0 too): the reference returns 0 early (atan2.rs:16-18)
    const int shift = __clz((int)x);
    y <<= shift;
    x <<= shift;
    constexpr int FRAC_BITS = 31 - IDSP_ATAN2_DIVI_DEPTH;
    const uint32_t rem = x & ((1u << FRAC_BITS) - 1);
    const uint32_t idx = (x << 1) >> (1 + FRAC_BITS);
    const uint2 e = SMEM_TAB ? tab[idx] : __ldg(tab + idx);
    // (slope as i64 * rem as i64) >> FRAC_BITS: rem < 2^27, so one signed IMAD.WIDE and a funnel shift
    const int64_t sp = mad_wide((int32_t)e.y, (int32_t)rem, 0);
    const uint32_t step = __funnelshift_r((uint32_t)sp, (uint32_t)((uint64_t)sp >> 32), FRAC_BITS);
    const uint32_t r0 = e.x + step;
    const uint32_t q = mul_q31(y, mul_q31(r0, 0u - mul_q31(x, r0)));
    return x_in ? q : 0u;
}
__device__ __forceinline__ uint32_t atani_dev(uint32_t x) {
    const int32_t ATANI[6] = {0x0517c2cd, -0x06c6496b, 0x0fbdb021,
                              -0x25b32e0a, 0x43b34c81, -0x3bc823dd};
    // ((x as i64 * x as i64) >> 32) as i32: the same bits as the unsigned high word
    const int32_t x2 = (int32_t)__umulhi(x, x);
    int32_t r = 0;
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        // Q32<32>*Q32<32> = (i64 product) >> 32, ops.rs:145-153: signed high word.  (mulhi_opaque: ptxas folds
        // mul.hi + add into IMAD.HI with a {0, coefficient} addend pair and rebuilds the pair with two moves
        // per Horner step: 13 moves per call on the multiplier pipe that bounds this function)
        r = mulhi_opaque(r, x2);
        r = (int32_t)((uint32_t)r + (uint32_t)ATANI[i]);
    }
    return (uint32_t)(((int64_t)r * (int64_t)(uint64_t)x) >> 28);
}
__device__ __forceinline__ int32_t sat_neg(int32_t a) {  // saturating_neg: 0 - a, saturated
    int32_t r;
    asm("sub.sat.s32 %0, 0, %1;" : "=r"(r) : "r"(a));
    return r;
}
template <bool SMEM_TAB>
__device__ __forceinline__ int32_t atan2_dev_t(const uint2 *tab, int32_t y, int32_t x) {
    uint32_t k = 0;
    if (y < 0) { y = sat_neg(y); k ^= 0xffffffffu; }
    if (x < 0) { x = sat_neg(x); k ^= 0xffffffffu >> 1; }
    if (y > x) { int32_t t = y; y = x; x = t; k ^= 0xffffffffu >> 2; }
    uint32_t r = atani_dev(divi_dev<SMEM_TAB>(tab, (uint32_t)y, (uint32_t)x));
    return (int32_t)(r ^ k);
}
__device__ __forceinline__ int32_t atan2_dev(int32_t y, int32_t x) {
    return atan2_dev_t<false>(g_divi_tab, y, x);
}

// --------------------------------------------------------------------------
// Lowpass<N> (src/lowpass.rs:47-78)
// --------------------------------------------------------------------------
__device__ __forceinline__ int32_t sat_sub(int32_t a, int32_t b) {
    int32_t r;
    asm("sub.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// IDSP_LP_DUPK (experiment, off): read the gains of the second `s1 += d` from a second copy in the kernel
// parameters so that ptxas cannot merge the two updates.  Measured effect: ptxas then emits 4 more
// IMAD.WIDE but keeps the 3-input carry-chain adds, so nothing is saved on the ALU pipe.
#ifdef IDSP_LP_DUPK
#define IDSP_KB(p, i) (p).kk[i]
#else
#define IDSP_KB(p, i) (p).k[i]
#endif
// k0b / k1b: the same gains read from a second copy in the kernel parameters.  ptxas cannot prove the two
// copies equal, so the second `s1 += d` stays two IMAD.WIDE (multiplier pipe, a quarter busy in the
// lock-in kernel) instead of being merged with the first into 64-bit carry-chain adds on the ALU pipe,
// which bounds that kernel: 4 IMAD.WIDE instead of 2 IMAD.WIDE + 4 IADD3 per step.
template <int ORDER, bool FAST = false>
__device__ __forceinline__ int32_t lowpass_step(int32_t k0, int32_t k1, int32_t k0b, int32_t k1b, int64_t &s0,
                                                int64_t &s1, int32_t x, uint32_t *flag = nullptr) {
    // d = dx*k0 (+ (s1>>32)*k1); every `+= d` below is folded into multiply-adds, which is
    // exact because i64 addition wraps (src/lowpass.rs:59-72 in release arithmetic)
    int32_t dx;
    if constexpr (FAST) {
        const int32_t sh = (int32_t)(s0 >> 32);
        dx = (int32_t)((uint32_t)x - (uint32_t)sh);
        *flag |= (uint32_t)sh ^ ((uint32_t)sh << 1);
    } else {
        dx = sat_sub(x, (int32_t)(s0 >> 32));
    }
    int32_t y;
    if constexpr (ORDER == 1) {
        s0 = mad_wide(dx, k0, s0);
        y = (int32_t)(s0 >> 32);
        s0 = mad_wide(dx, k0b, s0);
    } else {
        const int32_t s1h = (int32_t)(s1 >> 32);
        s1 = mad_wide(dx, k0, mad_wide(s1h, k1, s1));
        s0 = (int64_t)((uint64_t)s0 + (uint64_t)s1);
        y = (int32_t)(s0 >> 32);
        s0 = (int64_t)((uint64_t)s0 + (uint64_t)s1);
        s1 = mad_wide(dx, k0b, mad_wide(s1h, k1b, s1));
    }
    return y;
}
template <int ORDER> struct LowpassOp : OpHooks {
    using In = int32_t;
    using Out = int32_t;
    static constexpr bool LM_SMALL = true;
    struct Params {
        int32_t k[2];
        int32_t kk[2];  // = k (see lowpass_step)
        int64_t *st;
    };
    int64_t s0, s1;
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        s0 = p.st[lane];
        s1 = ORDER == 2 ? p.st[stride + lane] : 0;
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.st[lane] = s0;
        if (ORDER == 2) p.st[stride + lane] = s1;
    }
    __device__ __forceinline__ int32_t step(const Params &p, int32_t x) {
        return lowpass_step<ORDER>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), s0, s1, x);
    }
};

// Speculative saturation (tile kernels, tma_kernels.cuh).  `x.saturating_sub(s0 >> 32)` (src/lowpass.rs:56) costs
// five instructions (PTX sub.sat.s32 is emulated: subtract, two sign-logic predicates, two selects), four of them
// on the ALU pipe that bounds the lock-in kernels.  The mixer output is (lo * x) >> 32 with |lo| < 2^31 - 2^14
// (cossin's range), so |mix| < 2^30, and the difference cannot saturate while the state's high word lies in
// [-2^30, 2^30) -- i.e. while its two top bits agree.  step_fast() subtracts with wrap-around and ORs
// `sh ^ (sh << 1)` of every state word it used into a flag; a tile whose flag has bit 31 set is rolled back
// (spec_rollback) and redone with the exact step().  Results are bit-identical by construction; a lane whose
// low-pass state is within 6 dB of full scale pays the tile twice.  (336 -> 362 GSa/s on configs[3].)
#define IDSP_LOCKIN_SPEC_MEMBERS(RESTORE_EXTRA, SAVE_EXTRA)                                       \
    static constexpr bool SPECULATIVE = true;                                                      \
    uint32_t flag;                                                                                 \
    int64_t b_i0, b_i1, b_q0, b_q1;                                                                \
    __device__ __forceinline__ void spec_begin() {                                                 \
        flag = 0; b_i0 = i0; b_i1 = i1; b_q0 = q0; b_q1 = q1; SAVE_EXTRA                           \
    }                                                                                              \
    __device__ __forceinline__ bool spec_failed() const { return (int32_t)flag < 0; }              \
    __device__ __forceinline__ void spec_rollback() {                                              \
        i0 = b_i0; i1 = b_i1; q0 = b_q0; q1 = b_q1; RESTORE_EXTRA                                  \
    }
template <bool FAST, class Op, class P, class X> __device__ __forceinline__ auto op_step(Op &op, const P &p, X x) {
    if constexpr (FAST) return op.step_fast(p, x);
    else return op.step(p, x);
}
template <class Op, class = void> struct op_speculative { static constexpr bool value = false; };
template <class Op> struct op_speculative<Op, decltype((void)Op::SPECULATIVE)> { static constexpr bool value = Op::SPECULATIVE; };

// Accu (src/accu.rs:34-37) -> Complex::from_angle (src/complex.rs:237-240) ->
// Lockin<Lowpass<N>> (src/lockin.rs:17-39); mix = i32 * Q32<32> -> (lo*x)>>32
// (dsp-fixedpoint/src/lib.rs:449-456).
template <int ORDER, bool SMEM_LUT = false> struct LockinOp {
    using In = int32_t;
    using Out = int2;
    static constexpr bool TUNABLE = false;
    static constexpr bool HEAVY = true;
    static constexpr bool LM_SMALL = false;
    static constexpr int SMEM_EXTRA_WORDS = SMEM_LUT ? 256 * IDSP_LOCKIN_LUT_REP : 0;  // expanded cossin table staged per CTA
    struct Params {
        int32_t k[2];
        int32_t kk[2];  // = k (see lowpass_step)
        int32_t *accu_state;
        const int32_t *accu_step;
        int64_t *st;  // [2*ORDER][stride]
        const uint32_t *lut;
    };
    uint32_t ph, dph;
    int64_t i0, i1, q0, q1;
    const uint32_t *lutp;
    __device__ __forceinline__ static void init_smem(const Params &p, uint32_t *extra, int tid, int nthreads) {
        if constexpr (SMEM_LUT) cossin_expand_lut<IDSP_LOCKIN_LUT_REP>(p.lut, extra, tid, nthreads);
    }
    __device__ __forceinline__ void bind(const Params &p, const uint32_t *extra) {
        lutp = SMEM_LUT ? extra + 2 * (threadIdx.x % IDSP_LOCKIN_LUT_REP) : p.lut;
    }
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        if constexpr (!SMEM_LUT) lutp = p.lut;
        ph = (uint32_t)p.accu_state[lane];
        dph = (uint32_t)p.accu_step[lane];
        i0 = p.st[lane];
        i1 = ORDER == 2 ? p.st[stride + lane] : 0;
        q0 = p.st[(size_t)ORDER * stride + lane];
        q1 = ORDER == 2 ? p.st[(size_t)(ORDER + 1) * stride + lane] : 0;
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.accu_state[lane] = (int32_t)ph;
        p.st[lane] = i0;
        if (ORDER == 2) p.st[stride + lane] = i1;
        p.st[(size_t)ORDER * stride + lane] = q0;
        if (ORDER == 2) p.st[(size_t)(ORDER + 1) * stride + lane] = q1;
    }
    __device__ __forceinline__ int2 step(const Params &p, int32_t x) {
        ph += dph;
        int32_t c, s;
        if constexpr (SMEM_LUT) cossin_dev_x<IDSP_LOCKIN_LUT_REP>(lutp, (int32_t)ph, c, s);
        else cossin_dev<false>(lutp, (int32_t)ph, c, s);
        int32_t mi = (int32_t)(((int64_t)c * (int64_t)x) >> 32);
        int32_t mq = (int32_t)(((int64_t)s * (int64_t)x) >> 32);
        int2 r;
        r.x = lowpass_step<ORDER>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), i0, i1, mi);
        r.y = lowpass_step<ORDER>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), q0, q1, mq);
        return r;
    }
    IDSP_LOCKIN_SPEC_MEMBERS(ph = b_ph;, b_ph = ph;)
    uint32_t b_ph;
    __device__ __forceinline__ int2 step_fast(const Params &p, int32_t x) {
        ph += dph;
        int32_t c, s;
        if constexpr (SMEM_LUT) cossin_dev_x<IDSP_LOCKIN_LUT_REP, true>(lutp, (int32_t)ph, c, s);
        else cossin_dev<false>(lutp, (int32_t)ph, c, s);
        const int32_t mi = mulhi_opaque(c, x);
        const int32_t mq = mulhi_opaque(s, x);
        int2 r;
        r.x = lowpass_step<ORDER, true>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), i0, i1, mi, &flag);
        r.y = lowpass_step<ORDER, true>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), q0, q1, mq, &flag);
        return r;
    }
};

// Lockin<Lowpass<N>> on (sample, phase) tuples (src/lockin.rs:30-39): the phase comes with every sample
// instead of a per-lane Accu.  In = (x, phase).
template <int ORDER, bool SMEM_LUT = false> struct LockinPhaseOp {
    using In = int2;
    using Out = int2;
    static constexpr bool TUNABLE = false;
    static constexpr bool HEAVY = true;
    static constexpr bool LM_SMALL = false;
    static constexpr int SMEM_EXTRA_WORDS = SMEM_LUT ? 256 * IDSP_LOCKIN_LUT_REP : 0;
    struct Params {
        int32_t k[2];
        int32_t kk[2];  // = k (see lowpass_step)
        int64_t *st;  // [2*ORDER][stride]
        const uint32_t *lut;
    };
    int64_t i0, i1, q0, q1;
    const uint32_t *lutp;
    __device__ __forceinline__ static void init_smem(const Params &p, uint32_t *extra, int tid, int nthreads) {
        if constexpr (SMEM_LUT) cossin_expand_lut<IDSP_LOCKIN_LUT_REP>(p.lut, extra, tid, nthreads);
    }
    __device__ __forceinline__ void bind(const Params &p, const uint32_t *extra) {
        lutp = SMEM_LUT ? extra + 2 * (threadIdx.x % IDSP_LOCKIN_LUT_REP) : p.lut;
    }
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        if constexpr (!SMEM_LUT) lutp = p.lut;
        i0 = p.st[lane];
        i1 = ORDER == 2 ? p.st[stride + lane] : 0;
        q0 = p.st[(size_t)ORDER * stride + lane];
        q1 = ORDER == 2 ? p.st[(size_t)(ORDER + 1) * stride + lane] : 0;
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.st[lane] = i0;
        if (ORDER == 2) p.st[stride + lane] = i1;
        p.st[(size_t)ORDER * stride + lane] = q0;
        if (ORDER == 2) p.st[(size_t)(ORDER + 1) * stride + lane] = q1;
    }
    __device__ __forceinline__ int2 step(const Params &p, int2 xp) {
        int32_t c, s;
        if constexpr (SMEM_LUT) cossin_dev_x<IDSP_LOCKIN_LUT_REP>(lutp, xp.y, c, s);
        else cossin_dev<false>(lutp, xp.y, c, s);
        const int32_t mi = (int32_t)(((int64_t)c * (int64_t)xp.x) >> 32);
        const int32_t mq = (int32_t)(((int64_t)s * (int64_t)xp.x) >> 32);
        int2 r;
        r.x = lowpass_step<ORDER>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), i0, i1, mi);
        r.y = lowpass_step<ORDER>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), q0, q1, mq);
        return r;
    }
    IDSP_LOCKIN_SPEC_MEMBERS(, )
    __device__ __forceinline__ int2 step_fast(const Params &p, int2 xp) {
        int32_t c, s;
        if constexpr (SMEM_LUT) cossin_dev_x<IDSP_LOCKIN_LUT_REP, true>(lutp, xp.y, c, s);
        else cossin_dev<false>(lutp, xp.y, c, s);
        const int32_t mi = mulhi_opaque(c, xp.x);
        const int32_t mq = mulhi_opaque(s, xp.x);
        int2 r;
        r.x = lowpass_step<ORDER, true>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), i0, i1, mi, &flag);
        r.y = lowpass_step<ORDER, true>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), q0, q1, mq, &flag);
        return r;
    }
};

// Lockin<Lowpass<N>> on (sample, LO) tuples (src/lockin.rs:17-28), X = i32, U = Q32<32>.  In = (x, lo.re, lo.im).
struct XLo {
    int32_t x, re, im;
};
template <int ORDER> struct LockinLoOp : OpHooks {
    using In = XLo;
    using Out = int2;
    static constexpr bool HEAVY = true;
    struct Params {
        int32_t k[2];
        int32_t kk[2];  // = k (see lowpass_step)
        int64_t *st;
    };
    int64_t i0, i1, q0, q1;
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        i0 = p.st[lane];
        i1 = ORDER == 2 ? p.st[stride + lane] : 0;
        q0 = p.st[(size_t)ORDER * stride + lane];
        q1 = ORDER == 2 ? p.st[(size_t)(ORDER + 1) * stride + lane] : 0;
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.st[lane] = i0;
        if (ORDER == 2) p.st[stride + lane] = i1;
        p.st[(size_t)ORDER * stride + lane] = q0;
        if (ORDER == 2) p.st[(size_t)(ORDER + 1) * stride + lane] = q1;
    }
    __device__ __forceinline__ int2 step(const Params &p, XLo v) {
        const int32_t mi = (int32_t)(((int64_t)v.re * (int64_t)v.x) >> 32);
        const int32_t mq = (int32_t)(((int64_t)v.im * (int64_t)v.x) >> 32);
        int2 r;
        r.x = lowpass_step<ORDER>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), i0, i1, mi);
        r.y = lowpass_step<ORDER>(p.k[0], p.k[1], IDSP_KB(p, 0), IDSP_KB(p, 1), q0, q1, mq);
        return r;
    }
};

// --------------------------------------------------------------------------
// PLL (src/pll.rs:88-108): type-2 sampled-phase PLL, wrapping 32/64-bit integer math, with the
// ClampWrap phase-error clamp (src/unwrap.rs:166-194, overflowing_sub :73-81).  SURVEY 8(f) rank 4.
// State words (i32): [x0, clamp, z0, y0, f0 lo, f0 hi, f lo, f hi, y].
// --------------------------------------------------------------------------
struct PllOp : OpHooks {
    using In = int32_t;
    using Out = int32_t;
    static constexpr bool HEAVY = true;
    struct Params {
        int32_t ba[3];
        int32_t *st;
    };
    int32_t x0, clamp, z0, y0, y;
    int64_t f0, f;
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        x0 = p.st[lane];
        clamp = p.st[stride + lane];
        z0 = p.st[2 * stride + lane];
        y0 = p.st[3 * stride + lane];
        f0 = (int64_t)((uint64_t)(uint32_t)p.st[4 * stride + lane] | ((uint64_t)(uint32_t)p.st[5 * stride + lane] << 32));
        f = (int64_t)((uint64_t)(uint32_t)p.st[6 * stride + lane] | ((uint64_t)(uint32_t)p.st[7 * stride + lane] << 32));
        y = p.st[8 * stride + lane];
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.st[lane] = x0;
        p.st[stride + lane] = clamp;
        p.st[2 * stride + lane] = z0;
        p.st[3 * stride + lane] = y0;
        p.st[4 * stride + lane] = (int32_t)(uint32_t)f0;
        p.st[5 * stride + lane] = (int32_t)(f0 >> 32);
        p.st[6 * stride + lane] = (int32_t)(uint32_t)f;
        p.st[7 * stride + lane] = (int32_t)(f >> 32);
        p.st[8 * stride + lane] = y;
    }
    __device__ __forceinline__ int32_t step(const Params &p, int32_t x) {
        y = (int32_t)((uint32_t)y + (uint32_t)(int32_t)(f >> 32));  // oscillator, frequency() = f >> 32
        const int32_t t = (int32_t)((uint32_t)x + (uint32_t)y);     // phase error before the clamp
        const int32_t delta = (int32_t)((uint32_t)t - (uint32_t)x0);
        const int wrap = (int)(delta >= 0) - (int)(t >= x0);        // Ordering of the two bools
        x0 = t;
        const int c = clamp + wrap;
        clamp = (c > 0) - (c < 0);
        const int32_t o = clamp < 0 ? INT32_MIN : (clamp > 0 ? INT32_MAX : t);
        const int32_t zn = o >> 1;
        const int32_t yn = (int32_t)((uint32_t)zn + (uint32_t)z0);  // Nyquist zero
        z0 = zn;
        // lead-lag with a wide state: f0 += b0*y0 + b1*y0' + a1*hi(f0) + ((a1 * lo(f0)) >> 32)
        int64_t acc = mad_wide(p.ba[0], yn, f0);
        acc = mad_wide(p.ba[1], y0, acc);
        acc = mad_wide(p.ba[2], (int32_t)(f0 >> 32), acc);
        acc = (int64_t)((uint64_t)acc + (uint64_t)(((int64_t)p.ba[2] * (int64_t)(uint32_t)f0) >> 32));
        f0 = acc;
        y0 = yn;
        f = (int64_t)((uint64_t)f + (uint64_t)f0);  // DC pole
        return y;
    }
};

// --------------------------------------------------------------------------
// FM discriminator graph of examples/fm_disc.rs:26-48, fused:
//   z = x * prev.into_bits().conj()   Complex<Q32<32>> * Complex<i32> (src/complex.rs:117-134):
//                                     wide products, one late `>> 32` per component
//   d = z.arg() - carrier             atan2(im, re) (src/complex.rs:254-256), wrapping
//   y = Biquad<Q32<F>> DF1 (d)        src/iir/biquad.rs:366-383
// The first sample of a stream (no previous sample) gives d = 0.
// State words (i32): [has_prev, prev.re, prev.im, x1, x2, y1, y2].
// --------------------------------------------------------------------------
template <int MODE = 0>  // MODE 1: 0 <= F < 32 (funnel-shift quantiser, no test of F per sample)
struct FmDiscOp : OpHooks {
    using In = int2;   // Complex<Q32<32>> as raw (re, im)
    using Out = int32_t;
    static constexpr bool HEAVY = true;  // ~90 instructions per sample (atan2): per-warp pipelines
    struct Params {
        int32_t carrier;
        int32_t ba[5];
        int F;
        int32_t *st;
    };
    int32_t has, pre, pim, x1, x2, y1, y2;
    __device__ __forceinline__ void load(const Params &p, size_t lane, size_t stride) {
        has = p.st[lane];
        pre = p.st[stride + lane];
        pim = p.st[2 * stride + lane];
        x1 = p.st[3 * stride + lane];
        x2 = p.st[4 * stride + lane];
        y1 = p.st[5 * stride + lane];
        y2 = p.st[6 * stride + lane];
    }
    __device__ __forceinline__ void store(const Params &p, size_t lane, size_t stride) const {
        p.st[lane] = has;
        p.st[stride + lane] = pre;
        p.st[2 * stride + lane] = pim;
        p.st[3 * stride + lane] = x1;
        p.st[4 * stride + lane] = x2;
        p.st[5 * stride + lane] = y1;
        p.st[6 * stride + lane] = y2;
    }
    __device__ __forceinline__ int32_t step(const Params &p, int2 x) {
        // evaluated unconditionally and selected (no branch: the unrolled frame loop interleaves the
        // discriminators of several samples, only the biquad below is a recurrence)
        const int32_t cim = (int32_t)(0u - (uint32_t)pim);  // conj() of the i32 bits (wrapping neg)
        const int64_t re = (int64_t)((uint64_t)((int64_t)x.x * pre) - (uint64_t)((int64_t)x.y * cim));
        const int64_t im = (int64_t)((uint64_t)((int64_t)x.x * cim) + (uint64_t)((int64_t)x.y * pre));
        const int32_t dd = (int32_t)((uint32_t)atan2_dev((int32_t)(im >> 32), (int32_t)(re >> 32)) - (uint32_t)p.carrier);
        const int32_t d = has ? dd : 0;
        has = 1;
        pre = x.x;
        pim = x.y;
        const int32_t y0 = MODE == 1 ? SosI32Fast::eval(p.ba, p.F, d, x1, x2, y1, y2)
                                     : Sos<int32_t>::eval(p.ba, p.F, d, x1, x2, y1, y2);
        x2 = x1;
        x1 = d;
        y2 = y1;
        y1 = y0;
        return y0;
    }
};

}  // namespace idsp
