// biquad.cu -- C-ABI entry points of the iir::Biquad family (include/idsp_b200.h).
#include <string.h>

#include "common.cuh"
#include "ops.cuh"
#include "lane_kernels.cuh"
#include "tma_kernels.cuh"

using namespace idsp;

template <class T> static bool f_ok(int F) {
    if (is_float<T>::value) return true;
    return F > -(int)(8 * sizeof(T)) && F < (int)(16 * sizeof(T));
}

#define COMMON_ARGS_CHECK()                                                          \
    do {                                                                             \
        int r_ = idsp_use_device(ctx);                                               \
        if (r_) return r_;                                                           \
        IDSP_CHECK_ARG(ba != nullptr, "ba is null");                                 \
        IDSP_CHECK_ARG(layout == IDSP_FRAME_MAJOR || layout == IDSP_LANE_MAJOR,      \
                       "layout must be 0 (frame-major) or 1 (lane-major)");          \
        if (frames == 0 || lanes == 0) return IDSP_OK;                               \
        IDSP_CHECK_ARG(state != nullptr && x != nullptr && y != nullptr,             \
                       "state/x/y must not be null");                                \
    } while (0)

// ------------------------------------------------------------------ DF1
// i16 frame-major: 2 adjacent lanes share a 32-bit word, so the rows go through the frame-major TMA
// kernels as words (PackedOp, 794 -> 870 GSa/s); everything else takes the generic lane kernels
// (i8 packed four to a word: from 2^18 lanes on, see below).
template <class Op>
static int launch_packed_or_lanes(idsp_ctx *ctx, const typename Op::Params &p, const typename Op::In *x,
                                  typename Op::Out *y, size_t frames, size_t lanes, size_t sstride, int layout) {
    if constexpr (sizeof(typename Op::In) <= 2 && std::is_integral<typename Op::In>::value) {
        constexpr size_t P = 4 / sizeof(typename Op::In);
        // i8: four lanes per word quarter the number of threads, which only pays once there are enough lanes to
        // keep the SMs full: 65 536 lanes 852 -> 698 GSa/s, 2^18 lanes 1 095 -> 1 448, 2^20 lanes 1 149 -> 1 939, 2^22 lanes
        // 1 162 -> 2 013 (tools/bench_i8_packed.py); packed from 2^18 lanes on
        const char *e8 = getenv("IDSP_I8_PACKED_MIN_LANES");  // A/B switch, read per call (no shared mutable state)
        const size_t i8_min_lanes = e8 ? (size_t)atoll(e8) : (size_t)1 << 18;
        const bool enough = sizeof(typename Op::In) == 2 || lanes >= i8_min_lanes;
        if (enough && layout == IDSP_FRAME_MAJOR && lanes % (4 * P) == 0 && frames >= 16 &&
            (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
            int tr = tma_try_launch<PackedOp<Op>>(ctx, p, reinterpret_cast<const int32_t *>(x),
                                                  reinterpret_cast<int32_t *>(y), frames, lanes / P, sstride, layout);
            if (tr != IDSP_TMA_NOT_APPLICABLE) return tr;
        }
    }
    return launch_lanes_best<Op>(ctx, p, x, y, frames, lanes, sstride, layout);  // 8-byte samples: TMA kernels
}

template <class T>
static int df1_impl(idsp_ctx *ctx, const T *ba, int F, const T *clamp, T *state, const T *x,
                    T *y, size_t frames, size_t lanes, size_t sstride, int layout) {
    if (clamp) {
        typename Df1Op<T, true>::Params p;
        for (int i = 0; i < 5; i++) p.ba[i] = ba[i];
        p.F = F;
        p.u = clamp[0];
        p.mn = clamp[1];
        p.mx = clamp[2];
        p.st = state;
        return launch_packed_or_lanes<Df1Op<T, true>>(ctx, p, x, y, frames, lanes, sstride, layout);
    }
    typename Df1Op<T, false>::Params p;
    for (int i = 0; i < 5; i++) p.ba[i] = ba[i];
    p.F = F;
    p.u = p.mn = p.mx = T(0);
    p.st = state;
    return launch_packed_or_lanes<Df1Op<T, false>>(ctx, p, x, y, frames, lanes, sstride, layout);
}

// i32 specialisation: funnel-shift fast path for 0 <= F < 32 and the TMA-pipelined kernels
template <>
int df1_impl<int32_t>(idsp_ctx *ctx, const int32_t *ba, int F, const int32_t *clamp,
                      int32_t *state, const int32_t *x, int32_t *y, size_t frames, size_t lanes,
                      size_t sstride, int layout) {
    const bool fast = F >= 0 && F < 32;
#define GO(...)                                                                      \
    do {                                                                             \
        using OP = __VA_ARGS__;                                                      \
        typename OP::Params p;                                                       \
        for (int i = 0; i < 5; i++) p.ba[i] = ba[i];                                 \
        p.F = F;                                                                     \
        p.u = clamp ? clamp[0] : 0;                                                  \
        p.mn = clamp ? clamp[1] : 0;                                                 \
        p.mx = clamp ? clamp[2] : 0;                                                 \
        p.st = state;                                                                \
        int tr = tma_try_launch<OP>(ctx, p, x, y, frames, lanes, sstride, layout);   \
        if (tr != IDSP_TMA_NOT_APPLICABLE) return tr;                                \
        return launch_lanes<OP>(ctx, p, x, y, frames, lanes, sstride, layout);       \
    } while (0)
    if (clamp) {
        if (fast) GO(Df1Op<int32_t, true, 1>);
        GO(Df1Op<int32_t, true, 0>);
    }
    if (fast) GO(Df1Op<int32_t, false, 1>);
    GO(Df1Op<int32_t, false, 0>);
#undef GO
}

template <>
int df1_impl<float>(idsp_ctx *ctx, const float *ba, int F, const float *clamp, float *state,
                    const float *x, float *y, size_t frames, size_t lanes, size_t sstride,
                    int layout) {
#define GO(...)                                                                      \
    do {                                                                             \
        using OP = __VA_ARGS__;                                                      \
        typename OP::Params p;                                                       \
        for (int i = 0; i < 5; i++) p.ba[i] = ba[i];                                 \
        p.F = F;                                                                     \
        p.u = clamp ? clamp[0] : 0.f;                                                \
        p.mn = clamp ? clamp[1] : 0.f;                                               \
        p.mx = clamp ? clamp[2] : 0.f;                                               \
        p.st = state;                                                                \
        int tr = tma_try_launch<OP>(ctx, p, x, y, frames, lanes, sstride, layout);   \
        if (tr != IDSP_TMA_NOT_APPLICABLE) return tr;                                \
        return launch_lanes<OP>(ctx, p, x, y, frames, lanes, sstride, layout);       \
    } while (0)
    if (clamp) GO(Df1Op<float, true, 0>);
    GO(Df1Op<float, false, 0>);
#undef GO
}

int idsp_df1_f32_strided(idsp_ctx *ctx, const float *ba, float *state, const float *x, float *y, size_t frames,
                         size_t lanes, size_t sstride, int layout) {
    return df1_impl<float>(ctx, ba, 0, nullptr, state, x, y, frames, lanes, sstride, layout);
}

template <class T>
static int df1_host(idsp_ctx *ctx, const T *ba, int F, const T *clamp, T *state, const T *x, T *y,
                    size_t frames, size_t lanes, int layout) {
    HostStreamSpec s;
    s.frames = frames;
    s.lanes = lanes;
    s.in_bytes_per_frame_lane = sizeof(T);
    s.out_bytes_per_frame_lane = sizeof(T);
    s.layout = layout;
    s.nblobs = 1;
    s.blobs[0] = {state, 4 * lanes * sizeof(T), true};
    return idsp_host_stream(ctx, s, x, y,
                            [&](void **blobs, const void *dx, void *dy, size_t a0, size_t an) {
                                T *st = (T *)blobs[0];
                                if (layout == IDSP_FRAME_MAJOR)
                                    return df1_impl<T>(ctx, ba, F, clamp, st, (const T *)dx,
                                                       (T *)dy, an, lanes, lanes, layout);
                                return df1_impl<T>(ctx, ba, F, clamp, st + a0, (const T *)dx,
                                                   (T *)dy, frames, an, lanes, layout);
                            });
}

template <class T>
static int cascade_impl(idsp_ctx *ctx, const T *ba, int F, int nsec, T *state, const T *x, T *y,
                        size_t frames, size_t lanes, size_t sstride, int layout) {
#define GO(N)                                                                        \
    do {                                                                             \
        typename CascadeOp<T, N>::Params p;                                          \
        for (int s = 0; s < N; s++)                                                  \
            for (int i = 0; i < 5; i++) p.ba[s][i] = ba[5 * s + i];                    \
        p.F = F;                                                                     \
        p.nsec = nsec;                                                               \
        p.st = state;                                                                \
        if constexpr (std::is_same<T, int32_t>::value) {                             \
            /* i32: 0 <= F < 32 takes the funnel-shift ops on the TMA kernels, other F the generic ones */ \
            if (F >= 0 && F < 32) {                                                  \
                typename CascadeOp<T, N, 1>::Params q;                               \
                memcpy(&q, &p, sizeof(q));                                           \
                return launch_lanes_best<CascadeOp<T, N, 1>>(ctx, q, x, y, frames, lanes, sstride, layout); \
            }                                                                        \
            return launch_lanes<CascadeOp<T, N>>(ctx, p, x, y, frames, lanes, sstride, layout); \
        } else {                                                                     \
            return launch_lanes_best<CascadeOp<T, N>>(ctx, p, x, y, frames, lanes, sstride, layout); \
        }                                                                            \
    } while (0)
    switch (nsec) {
        case 1: GO(1);
        case 2: GO(2);
        case 3: GO(3);
        case 4: GO(4);
        case 5: GO(5);
        case 6: GO(6);
        case 7: GO(7);
        default: GO(8);
    }
#undef GO
}

#define DEF_DF1(S, T)                                                                \
    extern "C" int idsp_biquad_df1_##S(idsp_ctx *ctx, const T ba[5], int F,          \
                                       const T *clamp, T *state, const T *x, T *y,   \
                                       size_t frames, size_t lanes, int layout) {    \
        COMMON_ARGS_CHECK();                                                         \
        IDSP_CHECK_ARG(f_ok<T>(F), "F out of range for this sample type");           \
        return df1_impl<T>(ctx, ba, F, clamp, state, x, y, frames, lanes, lanes, layout); \
    }                                                                                \
    extern "C" int idsp_biquad_df1_##S##_host(idsp_ctx *ctx, const T ba[5], int F,   \
                                              const T *clamp, T *state, const T *x,  \
                                              T *y, size_t frames, size_t lanes,     \
                                              int layout) {                          \
        COMMON_ARGS_CHECK();                                                         \
        IDSP_CHECK_ARG(f_ok<T>(F), "F out of range for this sample type");           \
        return df1_host<T>(ctx, ba, F, clamp, state, x, y, frames, lanes, layout);   \
    }                                                                                \
    extern "C" int idsp_biquad_cascade_##S(idsp_ctx *ctx, const T *ba, int F, int nsec, \
                                           T *state, const T *x, T *y, size_t frames, \
                                           size_t lanes, int layout) {               \
        COMMON_ARGS_CHECK();                                                         \
        IDSP_CHECK_ARG(f_ok<T>(F), "F out of range for this sample type");           \
        IDSP_CHECK_ARG(nsec >= 1 && nsec <= IDSP_MAX_SECTIONS, "nsec out of range"); \
        return cascade_impl<T>(ctx, ba, F, nsec, state, x, y, frames, lanes, lanes, layout); \
    }
DEF_DF1(i8, int8_t)
DEF_DF1(i16, int16_t)
DEF_DF1(i32, int32_t)
DEF_DF1(i64, int64_t)
DEF_DF1(f32, float)
DEF_DF1(f64, double)

// ------------------------------------------------------------------ DF2T
template <class T>
static int df2t_impl(idsp_ctx *ctx, const T *ba, const T *clamp, T *state, const T *x, T *y,
                     size_t frames, size_t lanes, int layout) {
    COMMON_ARGS_CHECK();
    if (clamp) {
        typename Df2tOp<T, true>::Params p;
        for (int i = 0; i < 5; i++) p.ba[i] = ba[i];
        p.u = clamp[0];
        p.mn = clamp[1];
        p.mx = clamp[2];
        p.st = state;
        return launch_lanes_best<Df2tOp<T, true>>(ctx, p, x, y, frames, lanes, lanes, layout);
    }
    typename Df2tOp<T, false>::Params p;
    for (int i = 0; i < 5; i++) p.ba[i] = ba[i];
    p.u = p.mn = p.mx = T(0);
    p.st = state;
    return launch_lanes_best<Df2tOp<T, false>>(ctx, p, x, y, frames, lanes, lanes, layout);
}
extern "C" int idsp_biquad_df2t_f32(idsp_ctx *ctx, const float ba[5], const float *clamp,
                                    float *state, const float *x, float *y, size_t frames,
                                    size_t lanes, int layout) {
    return df2t_impl<float>(ctx, ba, clamp, state, x, y, frames, lanes, layout);
}
extern "C" int idsp_biquad_df2t_f64(idsp_ctx *ctx, const double ba[5], const double *clamp,
                                    double *state, const double *x, double *y, size_t frames,
                                    size_t lanes, int layout) {
    return df2t_impl<double>(ctx, ba, clamp, state, x, y, frames, lanes, layout);
}

// ------------------------------------------------------------------ Wide / Dither
template <template <bool> class OP>
static int i32_variant(idsp_ctx *ctx, const int32_t *ba, int F, const int32_t *clamp,
                       int32_t *state, const int32_t *x, int32_t *y, size_t frames, size_t lanes,
                       int layout) {
    COMMON_ARGS_CHECK();
    IDSP_CHECK_ARG(F >= 0 && F < 32, "F must satisfy 0 <= F < 32 (biquad.rs:458-460)");
    if (clamp) {
        typename OP<true>::Params p;
        for (int i = 0; i < 5; i++) p.ba[i] = ba[i];
        p.F = F;
        p.u = clamp[0];
        p.mn = clamp[1];
        p.mx = clamp[2];
        p.st = state;
        return launch_lanes_best<OP<true>>(ctx, p, x, y, frames, lanes, lanes, layout);
    }
    typename OP<false>::Params p;
    for (int i = 0; i < 5; i++) p.ba[i] = ba[i];
    p.F = F;
    p.u = p.mn = p.mx = 0;
    p.st = state;
    return launch_lanes_best<OP<false>>(ctx, p, x, y, frames, lanes, lanes, layout);
}
extern "C" int idsp_biquad_df1wide_i32(idsp_ctx *ctx, const int32_t ba[5], int F,
                                       const int32_t *clamp, int32_t *state, const int32_t *x,
                                       int32_t *y, size_t frames, size_t lanes, int layout) {
    return i32_variant<Df1WideOp>(ctx, ba, F, clamp, state, x, y, frames, lanes, layout);
}
extern "C" int idsp_biquad_df1dither_i32(idsp_ctx *ctx, const int32_t ba[5], int F,
                                         const int32_t *clamp, int32_t *state, const int32_t *x,
                                         int32_t *y, size_t frames, size_t lanes, int layout) {
    return i32_variant<Df1DitherOp>(ctx, ba, F, clamp, state, x, y, frames, lanes, layout);
}
