// hbf.cu -- C-ABI entry points of the half-band FIR family (include/idsp_b200.h).
#include <stdlib.h>

#include "common.cuh"
#include "hbf_stages.cuh"
#include "ops.cuh"
#include "hbf_fast_scalar.cuh"
#include "hbf_int_fast.cuh"

using namespace idsp;

static const float H_TAPS0[23] = IDSP_HBF_TAPS0;
static const float H_TAPS1[10] = IDSP_HBF_TAPS1;
static const float H_TAPS2[5] = IDSP_HBF_TAPS2;
static const float H_TAPS3[4] = IDSP_HBF_TAPS3;
static const float H_TAPS4[3] = IDSP_HBF_TAPS4;

extern "C" const float *idsp_hbf_taps(int index, int *M) {
    static const float *t[5] = {H_TAPS0, H_TAPS1, H_TAPS2, H_TAPS3, H_TAPS4};
    if (index < 0 || index > 4) return nullptr;
    if (M) *M = hbf_m(index);
    return t[index];
}
extern "C" size_t idsp_hbf_dec_state_words(int k) { return (k < 0 || k > 5) ? 0 : (size_t)hbf_dec_words(k); }
extern "C" size_t idsp_hbf_int_state_words(int k) { return (k < 0 || k > 5) ? 0 : (size_t)hbf_int_words(k); }
extern "C" size_t idsp_chain_state_words(int k) {
    return (k < 1 || k > 5) ? 0 : (size_t)(hbf_dec_words(k) + hbf_int_words(k) + 4);
}

// ---------------------------------------------------------------- register cascades
template <int K> struct DecCascadeRegs {
    DecStageRegs<K - 1> top;
    DecCascadeRegs<K - 1> rest;
    __device__ __forceinline__ void load(const float *st, size_t stride, size_t lane) {
        top.load(st, stride, lane);
        rest.load(st + (size_t)DecStageRegs<K - 1>::WORDS * stride, stride, lane);
    }
    __device__ __forceinline__ void store(float *st, size_t stride, size_t lane) const {
        top.store(st, stride, lane);
        rest.store(st + (size_t)DecStageRegs<K - 1>::WORDS * stride, stride, lane);
    }
    __device__ __forceinline__ float push(const float *x) {
        float half[1 << (K - 1)];
#pragma unroll
        for (int j = 0; j < (1 << (K - 1)); j++) half[j] = top.push(x[2 * j], x[2 * j + 1]);
        return rest.push(half);
    }
};
template <> struct DecCascadeRegs<0> {
    __device__ __forceinline__ void load(const float *, size_t, size_t) {}
    __device__ __forceinline__ void store(float *, size_t, size_t) const {}
    __device__ __forceinline__ float push(const float *x) { return x[0]; }
};
template <int K> struct IntCascadeRegs {
    IntCascadeRegs<K - 1> low;
    IntStageRegs<K - 1> top;
    __device__ __forceinline__ void load(const float *st, size_t stride, size_t lane) {
        low.load(st, stride, lane);
        top.load(st + (size_t)hbf_int_words(K - 1) * stride, stride, lane);
    }
    __device__ __forceinline__ void store(float *st, size_t stride, size_t lane) const {
        low.store(st, stride, lane);
        top.store(st + (size_t)hbf_int_words(K - 1) * stride, stride, lane);
    }
    __device__ __forceinline__ void push(float x, float *out) {
        float tmp[1 << (K - 1)];
        low.push(x, tmp);
#pragma unroll
        for (int j = 0; j < (1 << (K - 1)); j++) {
            float2 r = top.push(tmp[j]);
            out[2 * j] = r.x;
            out[2 * j + 1] = r.y;
        }
    }
};
template <> struct IntCascadeRegs<0> {
    __device__ __forceinline__ void load(const float *, size_t, size_t) {}
    __device__ __forceinline__ void store(float *, size_t, size_t) const {}
    __device__ __forceinline__ void push(float x, float *out) { out[0] = x; }
};

// vec != 0: the frame is 16-byte (R >= 4) / 8-byte (R == 2) aligned; otherwise scalar accesses
template <int R> __device__ __forceinline__ void load_frame(const float *p, float *v, int vec) {
    if (!vec) {
#pragma unroll
        for (int j = 0; j < R; j++) v[j] = p[j];
        return;
    }
    if constexpr (R >= 4) {
#pragma unroll
        for (int j = 0; j < R / 4; j++) {
            float4 q = reinterpret_cast<const float4 *>(p)[j];
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        }
    } else if constexpr (R == 2) {
        float2 q = *reinterpret_cast<const float2 *>(p);
        v[0] = q.x; v[1] = q.y;
    } else {
        v[0] = p[0];
    }
}
template <int R> __device__ __forceinline__ void store_frame(float *p, const float *v, int vec) {
    if (!vec) {
#pragma unroll
        for (int j = 0; j < R; j++) p[j] = v[j];
        return;
    }
    if constexpr (R >= 4) {
#pragma unroll
        for (int j = 0; j < R / 4; j++)
            reinterpret_cast<float4 *>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else if constexpr (R == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    } else {
        p[0] = v[0];
    }
}

// frame index helpers: element offset of (frame n, lane) in units of frames
__device__ __forceinline__ size_t fidx(int layout, size_t n, size_t lane, size_t nframes, size_t lanes) {
    return layout == IDSP_FRAME_MAJOR ? n * lanes + lane : lane * nframes + n;
}

// Generic thread-per-lane cascades (any shape / alignment). hbf_fast.cuh holds the
// shared-memory tiled kernels used for the large aligned cases.
// The delay lines live in registers as shift registers.  The main loops are unrolled over
// UNR frames (about 64 high-rate samples) so that the shifts inside the body become register
// renaming and only one real shift per delay line remains per iteration: with one frame per
// iteration the 23-tap stage alone would spend 65 MOVs per 70 flops.
__host__ __device__ constexpr int hbf_unroll(int K) { return K >= 6 ? 1 : (64 >> K) > 16 ? 16 : (64 >> K); }
template <int K>
__global__ void __launch_bounds__(128)
hbf_dec_cascade_generic(float *st, const float *x, float *y, size_t n_begin, size_t n_out,
                        size_t lanes, size_t sstride, int layout, int vec) {
    constexpr int R = 1 << K;
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    DecCascadeRegs<K> c;
    c.load(st, sstride, lane);
    constexpr int UNR = hbf_unroll(K);
    size_t n = n_begin;
    for (; n + UNR <= n_out; n += UNR) {
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            float v[R];
            size_t f = fidx(layout, n + u, lane, n_out, lanes);
            load_frame<R>(x + f * R, v, vec);
            y[f] = c.push(v);
        }
    }
    for (; n < n_out; n++) {
        float v[R];
        size_t f = fidx(layout, n, lane, n_out, lanes);
        load_frame<R>(x + f * R, v, vec);
        y[f] = c.push(v);
    }
    c.store(st, sstride, lane);
}
template <int K>
__global__ void __launch_bounds__(128)
hbf_int_cascade_generic(float *st, const float *x, float *y, size_t n_begin, size_t n_in,
                        size_t lanes, size_t sstride, int layout, int vec) {
    constexpr int R = 1 << K;
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    IntCascadeRegs<K> c;
    c.load(st, sstride, lane);
    constexpr int UNR = hbf_unroll(K);
    size_t n = n_begin;
    for (; n + UNR <= n_in; n += UNR) {
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            float v[R];
            size_t f = fidx(layout, n + u, lane, n_in, lanes);
            c.push(x[f], v);
            store_frame<R>(y + f * R, v, vec);
        }
    }
    for (; n < n_in; n++) {
        float v[R];
        size_t f = fidx(layout, n, lane, n_in, lanes);
        c.push(x[f], v);
        store_frame<R>(y + f * R, v, vec);
    }
    c.store(st, sstride, lane);
}
// config-5 chain: /2^K -> x2^K -> DF1 f32 (biquad.rs:366-383)
template <int K>
__global__ void __launch_bounds__(128)
chain_generic(float *st, Df1Op<float, false>::Params bp, const float *x, float *y, size_t n_low,
              size_t lanes, size_t sstride, int layout, int vec) {
    constexpr int R = 1 << K;
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    DecCascadeRegs<K> d;
    IntCascadeRegs<K> u_;
    Df1Op<float, false> b;
    d.load(st, sstride, lane);
    u_.load(st + (size_t)hbf_dec_words(K) * sstride, sstride, lane);
    bp.st = st + (size_t)(hbf_dec_words(K) + hbf_int_words(K)) * sstride;
    b.load(bp, lane, sstride);
    // (one frame per iteration: unrolling this body pushes it past 255 registers and is slower)
    for (size_t n = 0; n < n_low; n++) {
        float v[R], o[R];
        size_t f = fidx(layout, n, lane, n_low, lanes);
        load_frame<R>(x + f * R, v, vec);
        u_.push(d.push(v), o);
#pragma unroll
        for (int j = 0; j < R; j++) o[j] = b.step(bp, o[j]);
        store_frame<R>(y + f * R, o, vec);
    }
    d.store(st, sstride, lane);
    u_.store(st + (size_t)hbf_dec_words(K) * sstride, sstride, lane);
    b.store(bp, lane, sstride);
}

// ---------------------------------------------------------------- runtime-tap single stages
struct TapsParam {
    float c[IDSP_HBF_MAX_M];
    int M;
};
#define GEN_CH 8
// element offset (in frames of the stage) of frame `o` of `lane`: lane-major rows are contiguous; frame-major
// data may carry J frames of the stage per ABI frame and lane (the stages of a cascade: [f32; 2J] frames)
__device__ __forceinline__ size_t fidx2(int layout, size_t o, size_t lane, size_t n, size_t lanes, size_t J) {
    if (layout != IDSP_FRAME_MAJOR) return lane * n + o;
    return J == 1 ? o * lanes + lane : ((o / J) * lanes + lane) * J + (o % J);
}
// hbf.rs:46-68 window sum for runtime M
__device__ __forceinline__ float window_sum(const TapsParam &tp, const float *w, int odd, int sym) {
    const int M = tp.M;
    const float *nw = w + M + odd;
    float a0 = sym ? (nw[M - 1] + w[0]) : (nw[M - 1] - w[0]);
    float acc = a0 * tp.c[0];
    for (int i = 1; i < M; i++) {
        float a = sym ? (nw[M - 1 - i] + w[i]) : (nw[M - 1 - i] - w[i]);
        acc = acc + a * tp.c[i];
    }
    if (odd && sym) acc = acc + w[M];
    return acc;
}
// Any M <= IDSP_HBF_MAX_M: the windows are indexed dynamically (local memory).  Unaligned-safe: scalar accesses.
__global__ void __launch_bounds__(128)
hbf_dec_single(TapsParam tp, float *st, const float *x, float *y, size_t n_out, size_t lanes,
               size_t sstride, int layout, size_t J) {
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    const int M = tp.M, LEN = 2 * M - 1;
    float ev[IDSP_HBF_MAX_M + GEN_CH], od[2 * IDSP_HBF_MAX_M + GEN_CH];
    for (int i = 0; i < M - 1; i++) ev[i] = st[(size_t)i * sstride + lane];
    for (int i = 0; i < LEN; i++) od[i] = st[(size_t)(M - 1 + i) * sstride + lane];
    for (size_t o = 0; o < n_out; o += GEN_CH) {
        int c = (int)((n_out - o) < GEN_CH ? (n_out - o) : GEN_CH);
        for (int i = 0; i < c; i++) {
            const float *p = x + 2 * fidx2(layout, o + i, lane, n_out, lanes, J);
            ev[M - 1 + i] = p[0];
            od[LEN + i] = p[1];
        }
        for (int i = 0; i < c; i++)
            y[fidx2(layout, o + i, lane, n_out, lanes, J)] = window_sum(tp, od + i, 0, 1) + ev[i];
        for (int i = 0; i < M - 1; i++) ev[i] = ev[i + c];
        for (int i = 0; i < LEN; i++) od[i] = od[i + c];
    }
    for (int i = 0; i < M - 1; i++) st[(size_t)i * sstride + lane] = ev[i];
    for (int i = 0; i < LEN; i++) st[(size_t)(M - 1 + i) * sstride + lane] = od[i];
}
__global__ void __launch_bounds__(128)
hbf_int_single(TapsParam tp, float *st, const float *x, float *y, size_t n_in, size_t lanes,
               size_t sstride, int layout, size_t J) {
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    const int M = tp.M, LEN = 2 * M - 1;
    float xs[2 * IDSP_HBF_MAX_M + GEN_CH];
    for (int i = 0; i < LEN; i++) xs[i] = st[(size_t)i * sstride + lane];
    for (size_t o = 0; o < n_in; o += GEN_CH) {
        int c = (int)((n_in - o) < GEN_CH ? (n_in - o) : GEN_CH);
        for (int i = 0; i < c; i++) xs[LEN + i] = x[fidx2(layout, o + i, lane, n_in, lanes, J)];
        for (int i = 0; i < c; i++) {
            float *q = y + 2 * fidx2(layout, o + i, lane, n_in, lanes, J);
            q[0] = window_sum(tp, xs + i, 0, 1);
            q[1] = xs[M + i];
        }
        for (int i = 0; i < LEN; i++) xs[i] = xs[i + c];
    }
    for (int i = 0; i < LEN; i++) st[(size_t)i * sstride + lane] = xs[i];
}
// The tap counts of the reference's published designs (HBF_TAPS: 23, 10, 5, 4, 3; HBF_TAPS_98: 15, 6, 3, 3, 2)
// get the same kernels with M as a template parameter: every window index is static, so the delay
// lines live in registers (shifts become renaming inside the unrolled chunk) and the taps are read
// from the constant bank as instruction operands.
template <int M>
__global__ void __launch_bounds__(128)
hbf_dec_single_t(TapsParam tp, float *st, const float *x, float *y, size_t n_out, size_t lanes,
                 size_t sstride, int layout, size_t J, int vec) {
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    constexpr int LEN = 2 * M - 1, CH = M > 16 ? 4 : 8;
    float ev[M - 1 + CH], od[LEN + CH];
#pragma unroll
    for (int i = 0; i < M - 1; i++) ev[i] = st[(size_t)i * sstride + lane];
#pragma unroll
    for (int i = 0; i < LEN; i++) od[i] = st[(size_t)(M - 1 + i) * sstride + lane];
    size_t o = 0;
    for (; o + CH <= n_out; o += CH) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const float *p = x + 2 * fidx2(layout, o + i, lane, n_out, lanes, J);
            float2 v = vec ? *reinterpret_cast<const float2 *>(p) : make_float2(p[0], p[1]);
            ev[M - 1 + i] = v.x;
            od[LEN + i] = v.y;
        }
#pragma unroll
        for (int i = 0; i < CH; i++) {
            float acc = (od[i + LEN] + od[i]) * tp.c[0];
#pragma unroll
            for (int k = 1; k < M; k++) acc = acc + (od[i + LEN - k] + od[i + k]) * tp.c[k];
            y[fidx2(layout, o + i, lane, n_out, lanes, J)] = acc + ev[i];
        }
#pragma unroll
        for (int i = 0; i < M - 1; i++) ev[i] = ev[i + CH];
#pragma unroll
        for (int i = 0; i < LEN; i++) od[i] = od[i + CH];
    }
    for (; o < n_out; o++) {
        const float *p = x + 2 * fidx2(layout, o, lane, n_out, lanes, J);
        ev[M - 1] = p[0];
        od[LEN] = p[1];
        float acc = (od[LEN] + od[0]) * tp.c[0];
#pragma unroll
        for (int k = 1; k < M; k++) acc = acc + (od[LEN - k] + od[k]) * tp.c[k];
        y[fidx2(layout, o, lane, n_out, lanes, J)] = acc + ev[0];
#pragma unroll
        for (int i = 0; i < M - 1; i++) ev[i] = ev[i + 1];
#pragma unroll
        for (int i = 0; i < LEN; i++) od[i] = od[i + 1];
    }
#pragma unroll
    for (int i = 0; i < M - 1; i++) st[(size_t)i * sstride + lane] = ev[i];
#pragma unroll
    for (int i = 0; i < LEN; i++) st[(size_t)(M - 1 + i) * sstride + lane] = od[i];
}
template <int M>
__global__ void __launch_bounds__(128)
hbf_int_single_t(TapsParam tp, float *st, const float *x, float *y, size_t n_in, size_t lanes,
                 size_t sstride, int layout, size_t J, int vec) {
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    constexpr int LEN = 2 * M - 1, CH = M > 16 ? 4 : 8;
    float xs[LEN + CH];
#pragma unroll
    for (int i = 0; i < LEN; i++) xs[i] = st[(size_t)i * sstride + lane];
    size_t o = 0;
    for (; o + CH <= n_in; o += CH) {
#pragma unroll
        for (int i = 0; i < CH; i++) xs[LEN + i] = x[fidx2(layout, o + i, lane, n_in, lanes, J)];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            float acc = (xs[i + LEN] + xs[i]) * tp.c[0];
#pragma unroll
            for (int k = 1; k < M; k++) acc = acc + (xs[i + LEN - k] + xs[i + k]) * tp.c[k];
            float *q = y + 2 * fidx2(layout, o + i, lane, n_in, lanes, J);
            if (vec) *reinterpret_cast<float2 *>(q) = make_float2(acc, xs[M + i]);
            else { q[0] = acc; q[1] = xs[M + i]; }
        }
#pragma unroll
        for (int i = 0; i < LEN; i++) xs[i] = xs[i + CH];
    }
    for (; o < n_in; o++) {
        xs[LEN] = x[fidx2(layout, o, lane, n_in, lanes, J)];
        float acc = (xs[LEN] + xs[0]) * tp.c[0];
#pragma unroll
        for (int k = 1; k < M; k++) acc = acc + (xs[LEN - k] + xs[k]) * tp.c[k];
        float *q = y + 2 * fidx2(layout, o, lane, n_in, lanes, J);
        q[0] = acc;
        q[1] = xs[M];
#pragma unroll
        for (int i = 0; i < LEN; i++) xs[i] = xs[i + 1];
    }
#pragma unroll
    for (int i = 0; i < LEN; i++) st[(size_t)i * sstride + lane] = xs[i];
}
__global__ void __launch_bounds__(128)
fir_single(TapsParam tp, int odd, int sym, float *st, const float *x, float *y, size_t n,
           size_t lanes, size_t sstride, int layout) {
    size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    const int M = tp.M, LEN = 2 * M - 1 + odd;
    float xs[2 * IDSP_HBF_MAX_M + 1 + GEN_CH];
    for (int i = 0; i < LEN; i++) xs[i] = st[(size_t)i * sstride + lane];
    for (size_t o = 0; o < n; o += GEN_CH) {
        int c = (int)((n - o) < GEN_CH ? (n - o) : GEN_CH);
        for (int i = 0; i < c; i++) xs[LEN + i] = x[fidx(layout, o + i, lane, n, lanes)];
        for (int i = 0; i < c; i++)
            y[fidx(layout, o + i, lane, n, lanes)] = window_sum(tp, xs + i, odd, sym);
        for (int i = 0; i < LEN; i++) xs[i] = xs[i + c];
    }
    for (int i = 0; i < LEN; i++) st[(size_t)i * sstride + lane] = xs[i];
}

// ---------------------------------------------------------------- ABI
#define HBF_COMMON_CHECK(nframes)                                                    \
    do {                                                                             \
        int r_ = idsp_use_device(ctx);                                               \
        if (r_) return r_;                                                           \
        IDSP_CHECK_ARG(layout == IDSP_FRAME_MAJOR || layout == IDSP_LANE_MAJOR,      \
                       "layout must be 0 (frame-major) or 1 (lane-major)");          \
        if ((nframes) == 0 || lanes == 0) return IDSP_OK;                            \
        IDSP_CHECK_ARG(state != nullptr && x != nullptr && y != nullptr,             \
                       "state/x/y must not be null");                                \
    } while (0)

static int taps_param(const float *taps, int M, TapsParam *tp) {
    if (!taps || M < 1 || M > IDSP_HBF_MAX_M) {
        idsp_set_error("taps must be non-null and 1 <= M <= %d", IDSP_HBF_MAX_M);
        return IDSP_EINVAL;
    }
    for (int i = 0; i < IDSP_HBF_MAX_M; i++) tp->c[i] = i < M ? taps[i] : 0.f;
    tp->M = M;
    return IDSP_OK;
}

// one /2 or x2 stage on the ctx stream; J = frames of the stage per ABI frame and lane (frame-major cascades)
static int hbf_single_dev(idsp_ctx *ctx, bool dec, const TapsParam &tp, float *state, const float *x, float *y,
                          size_t n, size_t lanes, size_t sstride, int layout, size_t J) {
    const unsigned grid = (unsigned)((lanes + 127) / 128);
    const int vec = (((uintptr_t)(dec ? (const void *)x : (const void *)y)) & 7) == 0 ? 1 : 0;  // the pair side
#define GO(MM)                                                                                                   \
    case MM:                                                                                                     \
        if (dec) hbf_dec_single_t<MM><<<grid, 128, 0, ctx->stream>>>(tp, state, x, y, n, lanes, sstride, layout, J, vec); \
        else hbf_int_single_t<MM><<<grid, 128, 0, ctx->stream>>>(tp, state, x, y, n, lanes, sstride, layout, J, vec);     \
        IDSP_KERNEL_FAMILY(ctx, "hbf single stage, registers");                                                  \
        break;
    switch (ctx->policy == 1 ? 0 : tp.M) {
        GO(2) GO(3) GO(4) GO(5) GO(6) GO(10) GO(15) GO(23)
        default:
            if (dec) hbf_dec_single<<<grid, 128, 0, ctx->stream>>>(tp, state, x, y, n, lanes, sstride, layout, J);
            else hbf_int_single<<<grid, 128, 0, ctx->stream>>>(tp, state, x, y, n, lanes, sstride, layout, J);
            IDSP_KERNEL_FAMILY(ctx, "hbf single stage, run-time taps");
            break;
    }
#undef GO
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

extern "C" int idsp_hbf_dec_f32(idsp_ctx *ctx, const float *taps, int M, float *state,
                                const float *x, float *y, size_t n_out, size_t lanes, int layout) {
    HBF_COMMON_CHECK(n_out);
    TapsParam tp;
    int r = taps_param(taps, M, &tp);
    if (r) return r;
    return hbf_single_dev(ctx, true, tp, state, x, y, n_out, lanes, lanes, layout, 1);
}
extern "C" int idsp_hbf_int_f32(idsp_ctx *ctx, const float *taps, int M, float *state,
                                const float *x, float *y, size_t n_in, size_t lanes, int layout) {
    HBF_COMMON_CHECK(n_in);
    TapsParam tp;
    int r = taps_param(taps, M, &tp);
    if (r) return r;
    return hbf_single_dev(ctx, false, tp, state, x, y, n_in, lanes, lanes, layout, 1);
}
extern "C" int idsp_fir_f32(idsp_ctx *ctx, const float *taps, int M, int odd, int sym,
                            float *state, const float *x, float *y, size_t frames, size_t lanes,
                            int layout) {
    HBF_COMMON_CHECK(frames);
    TapsParam tp;
    int r = taps_param(taps, M, &tp);
    if (r) return r;
    fir_single<<<(unsigned)((lanes + 127) / 128), 128, 0, ctx->stream>>>(
        tp, odd ? 1 : 0, sym ? 1 : 0, state, x, y, frames, lanes, lanes, layout);
    IDSP_KERNEL_FAMILY(ctx, "fir single stage, run-time taps");
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

// frames of 2^k floats can be moved with 16-byte (k >= 2) / 8-byte (k == 1) accesses
static int frame_vec_ok(const void *p, int k) { return (((uintptr_t)p) & (k >= 2 ? 15 : 7)) == 0 ? 1 : 0; }

int hbf_dec_cascade_dev(idsp_ctx *ctx, int k, float *state, const float *x, float *y,
                        size_t n_out, size_t lanes, size_t sstride, int layout) {
    // large aligned lane-major streams: tiled TMA kernel over whole tiles, generic kernel
    // (state carried through `state`) for the remaining frames of every lane
    size_t done = 0;
    int fr = hbf_dec_fast_try_scalar(ctx, k, state, x, y, n_out, lanes, sstride, layout, &done);
    if (fr != IDSP_HBF_FAST_NOT_APPLICABLE && fr != IDSP_OK) return fr;
    if (fr == IDSP_HBF_FAST_NOT_APPLICABLE && ctx->policy == 2 && layout == IDSP_LANE_MAJOR) {
        idsp_set_error("tiled HBF kernel forced but shape/alignment does not qualify");
        return IDSP_EINVAL;
    }
    if (done == n_out) return IDSP_OK;
    const int vec = frame_vec_ok(x, k);  // the wide side: frames of 2^k floats
    unsigned grid = (unsigned)((lanes + 63) / 64);
    switch (k) {
        case 1: hbf_dec_cascade_generic<1><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_out, lanes, sstride, layout, vec); break;
        case 2: hbf_dec_cascade_generic<2><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_out, lanes, sstride, layout, vec); break;
        case 3: hbf_dec_cascade_generic<3><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_out, lanes, sstride, layout, vec); break;
        case 4: hbf_dec_cascade_generic<4><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_out, lanes, sstride, layout, vec); break;
        default: hbf_dec_cascade_generic<5><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_out, lanes, sstride, layout, vec); break;
    }
    IDSP_KERNEL_FAMILY(ctx, done ? "hbf tiled + generic tail" : "hbf generic thread-per-lane");
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

extern "C" int idsp_hbf_dec_cascade_f32(idsp_ctx *ctx, int log2_rate, float *state, const float *x,
                                        float *y, size_t n_out, size_t lanes, int layout) {
    HBF_COMMON_CHECK(n_out);
    IDSP_CHECK_ARG(log2_rate >= 1 && log2_rate <= 5, "log2_rate must be 1..5");
    return hbf_dec_cascade_dev(ctx, log2_rate, state, x, y, n_out, lanes, lanes, layout);
}
extern "C" int idsp_hbf_dec_cascade_f32_host(idsp_ctx *ctx, int log2_rate, float *state,
                                             const float *x, float *y, size_t n_out, size_t lanes,
                                             int layout) {
    HBF_COMMON_CHECK(n_out);
    IDSP_CHECK_ARG(log2_rate >= 1 && log2_rate <= 5, "log2_rate must be 1..5");
    HostStreamSpec s;
    s.frames = n_out;
    s.lanes = lanes;
    s.in_bytes_per_frame_lane = sizeof(float) << log2_rate;
    s.out_bytes_per_frame_lane = sizeof(float);
    s.layout = layout;
    s.nblobs = 1;
    s.blobs[0] = {state, (size_t)hbf_dec_words(log2_rate) * lanes * sizeof(float), true};
    return idsp_host_stream(ctx, s, x, y,
                            [&](void **blobs, const void *dx, void *dy, size_t a0, size_t an) {
                                float *st = (float *)blobs[0];
                                if (layout == IDSP_FRAME_MAJOR)
                                    return hbf_dec_cascade_dev(ctx, log2_rate, st, (const float *)dx,
                                                               (float *)dy, an, lanes, lanes, layout);
                                return hbf_dec_cascade_dev(ctx, log2_rate, st + a0, (const float *)dx,
                                                           (float *)dy, n_out, an, lanes, layout);
                            });
}
int hbf_int_cascade_dev(idsp_ctx *ctx, int log2_rate, float *state, const float *x, float *y, size_t n_in,
                        size_t lanes, size_t sstride, int layout) {
    size_t done = 0;
    int fr = hbf_int_fast_try(ctx, log2_rate, state, x, y, n_in, lanes, sstride, layout, &done);
    if (fr != IDSP_HBF_FAST_NOT_APPLICABLE && fr != IDSP_OK) return fr;
    if (fr == IDSP_HBF_FAST_NOT_APPLICABLE && ctx->policy == 2 && layout == IDSP_LANE_MAJOR) {
        idsp_set_error("tiled HBF kernel forced but shape/alignment does not qualify");
        return IDSP_EINVAL;
    }
    if (done == n_in) return IDSP_OK;
    const int vec = frame_vec_ok(y, log2_rate);
    unsigned grid = (unsigned)((lanes + 63) / 64);
    switch (log2_rate) {
        case 1: hbf_int_cascade_generic<1><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_in, lanes, sstride, layout, vec); break;
        case 2: hbf_int_cascade_generic<2><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_in, lanes, sstride, layout, vec); break;
        case 3: hbf_int_cascade_generic<3><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_in, lanes, sstride, layout, vec); break;
        case 4: hbf_int_cascade_generic<4><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_in, lanes, sstride, layout, vec); break;
        default: hbf_int_cascade_generic<5><<<grid, 64, 0, ctx->stream>>>(state, x, y, done, n_in, lanes, sstride, layout, vec); break;
    }
    IDSP_KERNEL_FAMILY(ctx, done ? "hbf tiled + generic tail" : "hbf generic thread-per-lane");
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}
extern "C" int idsp_hbf_int_cascade_f32(idsp_ctx *ctx, int log2_rate, float *state, const float *x,
                                        float *y, size_t n_in, size_t lanes, int layout) {
    HBF_COMMON_CHECK(n_in);
    IDSP_CHECK_ARG(log2_rate >= 1 && log2_rate <= 5, "log2_rate must be 1..5");
    return hbf_int_cascade_dev(ctx, log2_rate, state, x, y, n_in, lanes, lanes, layout);
}
// ---------------------------------------------------------------- caller-supplied tap sets
static const float H98_TAPS0[15] = {7.02144012e-05f, -2.43279582e-04f, 6.35026936e-04f, -1.39782541e-03f, 2.74613582e-03f,
                                    -4.96403839e-03f, 8.41806912e-03f, -1.35827601e-02f, 2.11004053e-02f, -3.19267647e-02f,
                                    4.77024289e-02f, -7.18014345e-02f, 1.12942004e-01f, -2.03279594e-01f, 6.33592923e-01f};
static const float H98_TAPS1[6] = {-0.00086943f, 0.00577837f, -0.02201674f, 0.06357869f, -0.16627679f, 0.61979312f};
static const float H98_TAPS2[3] = {0.01414651f, -0.10439639f, 0.59026742f};
static const float H98_TAPS3[3] = {0.01227974f, -0.09930782f, 0.58702834f};
static const float H98_TAPS4[2] = {-0.06291796f, 0.5629161f};
extern "C" const float *idsp_hbf_taps_98(int index, int *M) {  // src/hbf.rs:258-292
    static const float *t[5] = {H98_TAPS0, H98_TAPS1, H98_TAPS2, H98_TAPS3, H98_TAPS4};
    static const int m[5] = {15, 6, 3, 3, 2};
    if (index < 0 || index > 4) return nullptr;
    if (M) *M = m[index];
    return t[index];
}
extern "C" size_t idsp_hbf_cascade_state_words(int decimate, int nstages, const int *M) {
    if (!M || nstages < 1 || nstages > 5) return 0;
    size_t w = 0;
    for (int i = 0; i < nstages; i++) {
        if (M[i] < 1 || M[i] > IDSP_HBF_MAX_M) return 0;
        w += decimate ? (size_t)(3 * M[i] - 2) : (size_t)(2 * M[i] - 1);
    }
    return w;
}
// 1 = HBF_TAPS, 2 = HBF_TAPS_98 (compared by value: both sets are compiled into the tiled kernels), 0 = other
static int builtin_tap_set(int nstages, const float *const *taps, const int *M) {
    for (int set = 1; set <= 2; set++) {
        bool same = true;
        for (int i = 0; i < nstages && same; i++) {
            int m = 0;
            const float *t = set == 1 ? idsp_hbf_taps(i, &m) : idsp_hbf_taps_98(i, &m);
            same = M[i] == m && memcmp(taps[i], t, sizeof(float) * (size_t)m) == 0;
        }
        if (same) return set;
    }
    return 0;
}
static int cascade_taps_check(int nstages, const float *const *taps, const int *M) {
    if (nstages < 1 || nstages > 5 || !taps || !M) {
        idsp_set_error("nstages must be 1..5 and taps / M non-null");
        return IDSP_EINVAL;
    }
    for (int i = 0; i < nstages; i++)
        if (!taps[i] || M[i] < 1 || M[i] > IDSP_HBF_MAX_M) {
            idsp_set_error("stage %d: taps must be non-null and 1 <= M <= %d", i, IDSP_HBF_MAX_M);
            return IDSP_EINVAL;
        }
    return IDSP_OK;
}
// Decimator: stages nstages-1 -> 0 (src/hbf.rs:385-421).  Other tap sets than HBF_TAPS run stage by stage:
// stage s reads 2^(n-s) samples per output frame and lane and leaves 2^(n-s-1); the intermediate streams
// ping-pong through ctx scratch memory in the layout of x (frame-major: [t][lane][samples of the frame]).
extern "C" int idsp_hbf_dec_cascade_taps_f32(idsp_ctx *ctx, int nstages, const float *const *taps, const int *M,
                                             float *state, const float *x, float *y, size_t n_out, size_t lanes,
                                             int layout) {
    HBF_COMMON_CHECK(n_out);
    int r = cascade_taps_check(nstages, taps, M);
    if (r) return r;
    const int set = builtin_tap_set(nstages, taps, M);
    if (set == 1) return hbf_dec_cascade_dev(ctx, nstages, state, x, y, n_out, lanes, lanes, layout);
    // HBF_TAPS_98: the tiled kernels take the whole tiles of the call; a frame-major tail is contiguous and is
    // finished stage by stage below (the state is in the ABI layout in between), a lane-major call that is
    // not a whole number of tiles runs stage by stage as a whole
    if (set == 2 && (layout == IDSP_FRAME_MAJOR || n_out % ((size_t)hfs98::TT >> nstages) == 0)) {
        size_t done = 0;
        int fr = hbf98_dec_fast_try(ctx, nstages, state, x, y, n_out, lanes, lanes, layout, &done);
        if (fr != IDSP_HBF_FAST_NOT_APPLICABLE && fr != IDSP_OK) return fr;
        if (fr == IDSP_OK && done == n_out) return IDSP_OK;
        if (fr == IDSP_OK) {
            x += done * lanes << nstages;
            y += done * lanes;
            n_out -= done;
        }
    }
    const int n = nstages;
    void *scr = nullptr;
    const size_t half = n_out * lanes << (n - 1);  // floats after the first stage
    if (n > 1) {
        r = idsp_scratch(ctx, (half + half / 2) * sizeof(float), &scr);
        if (r) return r;
    }
    float *buf[2] = {(float *)scr, (float *)scr + half};
    const float *src = x;
    size_t word = 0;
    for (int s = 0; s < n; s++) {
        const int ti = n - 1 - s;
        TapsParam tp;
        r = taps_param(taps[ti], M[ti], &tp);
        if (r) return r;
        const size_t J = (size_t)1 << (n - 1 - s);  // stage outputs per ABI frame and lane
        float *dst = s == n - 1 ? y : buf[s & 1];
        r = hbf_single_dev(ctx, true, tp, state + word * lanes, src, dst, n_out * J, lanes, lanes, layout, J);
        if (r) return r;
        word += (size_t)(3 * M[ti] - 2);
        src = dst;
    }
    return IDSP_OK;
}
// Interpolator: stages 0 -> nstages-1 (src/hbf.rs:476-512)
extern "C" int idsp_hbf_int_cascade_taps_f32(idsp_ctx *ctx, int nstages, const float *const *taps, const int *M,
                                             float *state, const float *x, float *y, size_t n_in, size_t lanes,
                                             int layout) {
    HBF_COMMON_CHECK(n_in);
    int r = cascade_taps_check(nstages, taps, M);
    if (r) return r;
    const int set = builtin_tap_set(nstages, taps, M);
    if (set == 1) return idsp_hbf_int_cascade_f32(ctx, nstages, state, x, y, n_in, lanes, layout);
    if (set == 2 && (layout == IDSP_FRAME_MAJOR || n_in % ((size_t)hfi98::TOUT >> nstages) == 0)) {
        size_t done = 0;
        int fr = hbf98_int_fast_try(ctx, nstages, state, x, y, n_in, lanes, lanes, layout, &done);
        if (fr != IDSP_HBF_FAST_NOT_APPLICABLE && fr != IDSP_OK) return fr;
        if (fr == IDSP_OK && done == n_in) return IDSP_OK;
        if (fr == IDSP_OK) {
            x += done * lanes;
            y += done * lanes << nstages;
            n_in -= done;
        }
    }
    const int n = nstages;
    void *scr = nullptr;
    const size_t half = n_in * lanes << (n - 1);  // floats before the last stage
    if (n > 1) {
        r = idsp_scratch(ctx, (half + half / 2) * sizeof(float), &scr);
        if (r) return r;
    }
    // the largest intermediate (input of the last stage) sits in buf[(n-2) & 1]
    float *buf[2];
    buf[(n - 2) & 1] = (float *)scr;
    buf[((n - 2) & 1) ^ 1] = (float *)scr + half;
    const float *src = x;
    size_t word = 0;
    for (int s = 0; s < n; s++) {
        TapsParam tp;
        r = taps_param(taps[s], M[s], &tp);
        if (r) return r;
        const size_t J = (size_t)1 << s;  // stage inputs per ABI frame and lane
        float *dst = s == n - 1 ? y : buf[s & 1];
        r = hbf_single_dev(ctx, false, tp, state + word * lanes, src, dst, n_in * J, lanes, lanes, layout, J);
        if (r) return r;
        word += (size_t)(2 * M[s] - 1);
        src = dst;
    }
    return IDSP_OK;
}

// host buffers: the three operators in one PCIe round trip (lanes are independent: lane-major data is cut
// along lanes, frame-major along frames with the state carried from chunk to chunk)
extern "C" int idsp_chain_f32(idsp_ctx *ctx, int log2_rate, const float ba[5], float *state,
                              const float *x, float *y, size_t n_low, size_t lanes, int layout);
static int chain_dev_strided(idsp_ctx *ctx, int log2_rate, const float ba[5], float *state, const float *x, float *y,
                             size_t n_low, size_t lanes, size_t sstride, int layout);
extern "C" int idsp_chain_f32_host(idsp_ctx *ctx, int log2_rate, const float ba[5], float *state,
                                   const float *x, float *y, size_t n_low, size_t lanes, int layout) {
    HBF_COMMON_CHECK(n_low);
    IDSP_CHECK_ARG(ba != nullptr, "ba is null");
    IDSP_CHECK_ARG(log2_rate >= 1 && log2_rate <= 5, "log2_rate must be 1..5");
    HostStreamSpec s;
    s.frames = n_low;
    s.lanes = lanes;
    s.in_bytes_per_frame_lane = sizeof(float) << log2_rate;
    s.out_bytes_per_frame_lane = sizeof(float) << log2_rate;
    s.layout = layout;
    s.nblobs = 1;
    s.blobs[0] = {state, idsp_chain_state_words(log2_rate) * lanes * sizeof(float), true};
    return idsp_host_stream(ctx, s, x, y, [&](void **blobs, const void *dx, void *dy, size_t a0, size_t an) {
        float *st = (float *)blobs[0];
        if (layout == IDSP_FRAME_MAJOR)
            return chain_dev_strided(ctx, log2_rate, ba, st, (const float *)dx, (float *)dy, an, lanes, lanes, layout);
        return chain_dev_strided(ctx, log2_rate, ba, st + a0, (const float *)dx, (float *)dy, n_low, an, lanes, layout);
    });
}

extern "C" int idsp_chain_f32(idsp_ctx *ctx, int log2_rate, const float ba[5], float *state,
                              const float *x, float *y, size_t n_low, size_t lanes, int layout) {
    HBF_COMMON_CHECK(n_low);
    IDSP_CHECK_ARG(ba != nullptr, "ba is null");
    IDSP_CHECK_ARG(log2_rate >= 1 && log2_rate <= 5, "log2_rate must be 1..5");
    return chain_dev_strided(ctx, log2_rate, ba, state, x, y, n_low, lanes, lanes, layout);
}
// state words of lane l at state[w * sstride + l] (sstride >= lanes: a lane block of a larger SoA state)
static int chain_dev_strided(idsp_ctx *ctx, int log2_rate, const float ba[5], float *state, const float *x, float *y,
                             size_t n_low, size_t lanes, size_t sstride, int layout) {
    // Few lanes, lane-major: the fused thread-per-lane kernel would leave most SMs idle (one thread
    // per lane, ~200 registers of delay lines).  Run the three bit-identical pieces instead: the two
    // FIR cascades are time-parallel (tiled kernels, 8 lanes per CTA), only the biquad recurrence is
    // serial per lane.  The low-rate stream goes through ctx scratch, the biquad runs in place on y.
    // Measured cross-over on B200: composed wins up to 2^14 lanes (profiles/r1_bench_chain.json).
    // Lane-major, whole tiles, up to 2^18 lanes: tiled decimator -> low-rate scratch -> tiled interpolator
    // with the biquad fused as a fifth warp (hbf_int_fast_body.cuh): two passes, 4.25 + 4.25 bytes per sample.
    // Measured on B200 against the alternatives (GSa/s at 2^10 / 2^12 / 2^14 / 2^16 / 2^18 lanes, 2^30 samples):
    // three kernels 44 / 145 / 276 / - / -, thread-per-lane fused - / - / - / 320 / 353, this path with 8-lane
    // tiles 86 / 243 / 250 / 271 / 245 and with 16-lane tiles 71 / 221 / 313 / 337 / 304; since the biquad warp's
    // store wait left its critical path (round 2) 16-lane tiles give 335 / 364 / 352 / 340 at 2^14 / 2^16 / 2^17 /
    // 2^18 lanes against 322 for the single-pass kernel at 2^18 (tools/chain_crossover.py).  Short streams
    // (fewer than 8 tiles per call) and more lanes stay on the single-pass thread-per-lane kernel.
    int wide = lanes > 8192 ? 1 : 0;  // 0: 8-lane tiles, 1: 16-lane tiles
#ifdef IDSP_TUNE
    if (const char *e = getenv("IDSP_CHAIN_WIDE")) wide = atoi(e);  // 2: 32-lane tiles with 8 FIR warps
#endif
    const size_t tile_low = hbf_int_bq_tile(log2_rate, wide);
    if (layout == IDSP_LANE_MAJOR && ctx->policy != 1 && n_low % tile_low == 0 && n_low >= 8 * tile_low &&
        (lanes <= 262144 || ctx->policy == 2) && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0) {
        void *low = nullptr;
        int r = idsp_scratch(ctx, n_low * lanes * sizeof(float), &low);
        if (r) return r;
        const size_t wd = (size_t)hbf_dec_words(log2_rate), wi = (size_t)hbf_int_words(log2_rate);
        r = hbf_dec_cascade_dev(ctx, log2_rate, state, x, (float *)low, n_low, lanes, sstride, layout);
        if (r) return r;
        Df1Op<float, false>::Params bq;
        for (int i = 0; i < 5; i++) bq.ba[i] = ba[i];
        bq.F = 0;
        bq.u = bq.mn = bq.mx = 0.f;
        bq.st = state + (wd + wi) * sstride;
        r = hbf_int_bq_fast_try(ctx, log2_rate, state + wd * sstride, (const float *)low, y, n_low, lanes, sstride, bq, wide);
        if (r != IDSP_HBF_FAST_NOT_APPLICABLE) return r;
        // (not reached: the scratch buffer is aligned) the three-kernel composition
        r = hbf_int_cascade_dev(ctx, log2_rate, state + wd * sstride, (const float *)low, y, n_low, lanes, sstride, layout);
        if (r) return r;
        return idsp_df1_f32_strided(ctx, ba, state + (wd + wi) * sstride, y, y, n_low << log2_rate, lanes, sstride, layout);
    }
    if (layout == IDSP_LANE_MAJOR && ctx->policy != 1 && lanes <= 16384) {
        void *low = nullptr;
        int r = idsp_scratch(ctx, n_low * lanes * sizeof(float), &low);
        if (r) return r;
        const size_t wd = (size_t)hbf_dec_words(log2_rate), wi = (size_t)hbf_int_words(log2_rate);
        r = hbf_dec_cascade_dev(ctx, log2_rate, state, x, (float *)low, n_low, lanes, sstride, layout);
        if (r) return r;
        r = hbf_int_cascade_dev(ctx, log2_rate, state + wd * sstride, (const float *)low, y, n_low, lanes, sstride, layout);
        if (r) return r;
        return idsp_df1_f32_strided(ctx, ba, state + (wd + wi) * sstride, y, y, n_low << log2_rate, lanes, sstride, layout);
    }
    Df1Op<float, false>::Params bp;
    for (int i = 0; i < 5; i++) bp.ba[i] = ba[i];
    bp.F = 0;
    bp.u = bp.mn = bp.mx = 0.f;
    bp.st = nullptr;
    const int vec = frame_vec_ok(x, log2_rate) && frame_vec_ok(y, log2_rate);
    unsigned grid = (unsigned)((lanes + 63) / 64);
    switch (log2_rate) {
        case 1: chain_generic<1><<<grid, 64, 0, ctx->stream>>>(state, bp, x, y, n_low, lanes, sstride, layout, vec); break;
        case 2: chain_generic<2><<<grid, 64, 0, ctx->stream>>>(state, bp, x, y, n_low, lanes, sstride, layout, vec); break;
        case 3: chain_generic<3><<<grid, 64, 0, ctx->stream>>>(state, bp, x, y, n_low, lanes, sstride, layout, vec); break;
        case 4: chain_generic<4><<<grid, 64, 0, ctx->stream>>>(state, bp, x, y, n_low, lanes, sstride, layout, vec); break;
        default: chain_generic<5><<<grid, 64, 0, ctx->stream>>>(state, bp, x, y, n_low, lanes, sstride, layout, vec); break;
    }
    IDSP_KERNEL_FAMILY(ctx, "chain single pass thread-per-lane");
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}
