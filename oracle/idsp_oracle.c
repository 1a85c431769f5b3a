/*
 * idsp_oracle.c -- CPU restatement of the quartiq/idsp filter hot path.
 * TEST INFRASTRUCTURE ONLY (see idsp_oracle.h).  Compile with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp
 * Every function cites the reference file:line (relative to the reference tree).
 */
#include "idsp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef __int128 i128;
typedef unsigned __int128 u128;

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* Rust-semantics helpers (SURVEY.md 8a')                               */
/* ------------------------------------------------------------------ */
/* f64 -> iN `as` cast: saturating, NaN -> 0 */
static int64_t f64_as_i64(double v) {
    if (v != v) return 0;
    if (v >= 9223372036854775808.0) return INT64_MAX;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)v;
}
static int64_t sat_range(double v, int64_t lo, int64_t hi) {
    if (v != v) return 0;
    if (v >= (double)hi) return hi;
    if (v <= (double)lo) return lo;
    return (int64_t)v;
}
/* float -> Q: (v * 2^F).round() as T  (dsp-fixedpoint/src/num_traits_impl.rs:32-45)
 * round() = half away from zero */
int8_t orc_quant_i8(double v, int F) { return (int8_t)sat_range(round(v * ldexp(1.0, F)), INT8_MIN, INT8_MAX); }
int16_t orc_quant_i16(double v, int F) { return (int16_t)sat_range(round(v * ldexp(1.0, F)), INT16_MIN, INT16_MAX); }
int32_t orc_quant_i32(double v, int F) { return (int32_t)sat_range(round(v * ldexp(1.0, F)), INT32_MIN, INT32_MAX); }
int64_t orc_quant_i64(double v, int F) { return f64_as_i64(round(v * ldexp(1.0, F))); }
int32_t orc_round_sat_i32(double v) { return (int32_t)sat_range(round(v), INT32_MIN, INT32_MAX); }

/* src/iir/biquad.rs:547-562 */
void orc_ba_normalize_f64(const double ba6[6], double out5[5]) {
    double a0 = 1.0 / ba6[3];
    out5[0] = ba6[0] * a0;
    out5[1] = ba6[1] * a0;
    out5[2] = ba6[2] * a0;
    out5[3] = -ba6[4] * a0;
    out5[4] = -ba6[5] * a0;
}

static inline int32_t sat_sub_i32(int32_t a, int32_t b) {
    int64_t r = (int64_t)a - (int64_t)b;
    return r > INT32_MAX ? INT32_MAX : (r < INT32_MIN ? INT32_MIN : (int32_t)r);
}
static inline int32_t sat_neg_i32(int32_t a) { return a == INT32_MIN ? INT32_MAX : -a; }

/* ------------------------------------------------------------------ */
/* tables: build.rs:9-69                                               */
/* ------------------------------------------------------------------ */
static uint32_t g_cossin[128];
static uint32_t g_divi_base[16];
static int32_t g_divi_slope[16];
static int g_tables_ready = 0;

static void make_tables(void) {
    if (g_tables_ready) return;
    /* build.rs:28-41 */
    for (int i = 0; i < 128; i++) {
        double a = M_PI / 4. * (((double)i + 0.5) / 128.0);
        double s = sin(a), c = cos(a);
        uint32_t ci = (uint32_t)round((c * 2. - 1.) * 65535.0 - 1.);
        uint32_t si = (uint32_t)round(s * 65535.0);
        g_cossin[i] = ci + (si << 16);
    }
    /* build.rs:59-65 */
    const double Q31 = 2147483648.0;
    for (int i = 0; i < 16; i++) {
        double x0 = 1.0 + (double)i / 16.0;
        double x1 = 1.0 + (double)(i + 1) / 16.0;
        g_divi_base[i] = (uint32_t)round(Q31 / x0);
        g_divi_slope[i] = (int32_t)round((1.0 / x1 - 1.0 / x0) * Q31);
    }
    g_tables_ready = 1;
}
const uint32_t *orc_cossin_table(void) { make_tables(); return g_cossin; }
const uint32_t *orc_atan2_divi_base(void) { make_tables(); return g_divi_base; }
const int32_t *orc_atan2_divi_slope(void) { make_tables(); return g_divi_slope; }

/* ------------------------------------------------------------------ */
/* cossin: src/cossin.rs:14-67                                         */
/* ------------------------------------------------------------------ */
static inline void cossin_tab(const uint32_t *lut, int32_t phase, int32_t *co, int32_t *so) {
    uint32_t octant = (uint32_t)phase;
    if (octant & (1u << 29)) phase = ~phase;
    /* ALIGN_MSB = 15, COSSIN_DEPTH = 7: (phase<<3) >> (32-7-15) */
    phase = (int32_t)((((uint32_t)phase) << 3) >> 10);
    uint32_t lookup = lut[phase >> 15];
    phase &= (1 << 15) - 1;
    phase -= 1 << 14;
    const int32_t PI4 = 51471; /* (FRAC_PI_4 * 65536) as i32 */
    int32_t dphi = (phase * PI4) >> 16;
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
    *co = c;
    *so = s;
}
void orc_cossin(int32_t phase, int32_t *co, int32_t *so) {
    make_tables();
    cossin_tab(g_cossin, phase, co, so);
}
void orc_cossin_n(const int32_t *phase, int32_t *cs, size_t n) {
    make_tables();
    for (size_t i = 0; i < n; i++) cossin_tab(g_cossin, phase[i], &cs[2 * i], &cs[2 * i + 1]);
}

/* ------------------------------------------------------------------ */
/* atan2: src/atan2.rs:7-82                                            */
/* ------------------------------------------------------------------ */
static inline uint32_t mul_q31(uint32_t x, uint32_t y) { return (uint32_t)(((uint64_t)x * (uint64_t)y) >> 31); }

static inline uint32_t divi(uint32_t y, uint32_t x) {
    if (x == 0) return 0;
    int shift = __builtin_clz(x);
    y <<= shift;
    x <<= shift;
    const int FRAC_BITS = 31 - 4;
    uint32_t rem = x & ((1u << FRAC_BITS) - 1);
    uint32_t idx = (x << 1) >> (1 + FRAC_BITS);
    uint32_t base = g_divi_base[idx];
    int32_t slope = g_divi_slope[idx];
    uint32_t step = (uint32_t)(((int64_t)slope * (int64_t)rem) >> FRAC_BITS);
    uint32_t r0 = base + step;
    return mul_q31(y, mul_q31(r0, (uint32_t)(0u - mul_q31(x, r0))));
}

static inline uint32_t atani(uint32_t x) {
    static const int32_t ATANI[6] = {0x0517c2cd, -0x06c6496b, 0x0fbdb021,
                                     -0x25b32e0a, 0x43b34c81, -0x3bc823dd};
    /* x2 = ((x as i64 * x as i64) >> 32) as i32 */
    int32_t x2 = (int32_t)(((int64_t)(uint64_t)x * (int64_t)(uint64_t)x) >> 32);
    int32_t r = 0;
    for (int i = 5; i >= 0; i--) {
        /* Q32<32> * Q32<32> (dsp-fixedpoint/src/ops.rs:145-153) then + a */
        r = (int32_t)(((int64_t)r * (int64_t)x2) >> 32);
        r = (int32_t)((uint32_t)r + (uint32_t)ATANI[i]);
    }
    return (uint32_t)(((int64_t)r * (int64_t)(uint64_t)x) >> 28);
}

int32_t orc_atan2(int32_t y, int32_t x) {
    make_tables();
    uint32_t k = 0;
    if (y < 0) { y = sat_neg_i32(y); k ^= 0xffffffffu; }
    if (x < 0) { x = sat_neg_i32(x); k ^= 0xffffffffu >> 1; }
    if (y > x) { int32_t t = y; y = x; x = t; k ^= 0xffffffffu >> 2; }
    uint32_t r = atani(divi((uint32_t)y, (uint32_t)x));
    return (int32_t)(r ^ k);
}
void orc_atan2_n(const int32_t *xy, int32_t *p, size_t n) {
    for (size_t i = 0; i < n; i++) p[i] = orc_atan2(xy[2 * i + 1], xy[2 * i]);
}

/* ------------------------------------------------------------------ */
/* generic lanes driver                                                */
/* ------------------------------------------------------------------ */
/* Runs BODY(lane_lo, lane_hi) over contiguous lane blocks on nthreads. */
#define LANE_BLOCKS(lanes, nthreads, lo, hi, ...)                                   \
    do {                                                                             \
        int nt_ = (nthreads) < 1 ? 1 : (nthreads);                                   \
        if ((size_t)nt_ > (lanes)) nt_ = (int)((lanes) ? (lanes) : 1);               \
        _Pragma("omp parallel for num_threads(nt_) schedule(static)")                \
        for (int tb_ = 0; tb_ < nt_; tb_++) {                                        \
            size_t lo = (lanes) * (size_t)tb_ / (size_t)nt_;                         \
            size_t hi = (lanes) * (size_t)(tb_ + 1) / (size_t)nt_;                   \
            __VA_ARGS__                                                              \
        }                                                                            \
    } while (0)

/* ------------------------------------------------------------------ */
/* Biquad DF1 fixed point: src/iir/biquad.rs:366-383, Q arithmetic      */
/* dsp-fixedpoint/src/ops.rs:91-97, lib.rs:297-312, num_traits_impl.rs:74-104 */
/* ------------------------------------------------------------------ */
#define DEF_DF1_FIXED(S, T, UT, A, UA)                                               \
    static inline T df1_step_##S(const T *ba, int F, T *x1, T *x2, T *y1, T *y2,     \
                                 T x0) {                                             \
        UA acc = (UA)((A)ba[0] * (A)x0) + (UA)((A)ba[1] * (A)*x1) +                  \
                 (UA)((A)ba[2] * (A)*x2) + (UA)((A)ba[3] * (A)*y1) +                 \
                 (UA)((A)ba[4] * (A)*y2);                                            \
        A q = F >= 0 ? (A)((A)acc >> F) : (A)(UA)(acc << (-F));                      \
        T y0 = (T)(UT)(UA)q;                                                         \
        *x2 = *x1;                                                                   \
        *x1 = x0;                                                                    \
        *y2 = *y1;                                                                   \
        *y1 = y0;                                                                    \
        return y0;                                                                   \
    }                                                                                \
    static inline T df1c_step_##S(const T *ba, int F, const T *cl, T *x1, T *x2,     \
                                  T *y1, T *y2, T x0) {                              \
        T r = df1_step_##S(ba, F, x1, x2, y1, y2, x0);                               \
        if (cl) { /* biquad.rs:399-402 */                                            \
            T v = (T)(UT)((UT)r + (UT)cl[0]);                                        \
            v = v < cl[1] ? cl[1] : (v > cl[2] ? cl[2] : v);                         \
            *y1 = v;                                                                 \
            return v;                                                                \
        }                                                                            \
        return r;                                                                    \
    }

#define DEF_DF1_FLOAT(S, T)                                                          \
    static inline T df1_step_##S(const T *ba, int F, T *x1, T *x2, T *y1, T *y2,     \
                                 T x0) {                                             \
        (void)F;                                                                     \
        T y0 = ba[0] * x0 + ba[1] * *x1 + ba[2] * *x2 + ba[3] * *y1 + ba[4] * *y2;   \
        *x2 = *x1;                                                                   \
        *x1 = x0;                                                                    \
        *y2 = *y1;                                                                   \
        *y1 = y0;                                                                    \
        return y0;                                                                   \
    }                                                                                \
    static inline T df1c_step_##S(const T *ba, int F, const T *cl, T *x1, T *x2,     \
                                  T *y1, T *y2, T x0) {                              \
        T r = df1_step_##S(ba, F, x1, x2, y1, y2, x0);                               \
        if (cl) {                                                                    \
            T v = r + cl[0];                                                         \
            v = v < cl[1] ? cl[1] : (v > cl[2] ? cl[2] : v);                         \
            *y1 = v;                                                                 \
            return v;                                                                \
        }                                                                            \
        return r;                                                                    \
    }

DEF_DF1_FIXED(i8, int8_t, uint8_t, int16_t, uint16_t)
DEF_DF1_FIXED(i16, int16_t, uint16_t, int32_t, uint32_t)
DEF_DF1_FIXED(i32, int32_t, uint32_t, int64_t, uint64_t)
DEF_DF1_FIXED(i64, int64_t, uint64_t, i128, u128)
DEF_DF1_FLOAT(f32, float)
DEF_DF1_FLOAT(f64, double)

/* Public single-lane + lanes + cascade for each type */
#define DEF_DF1_API(S, T)                                                            \
    void orc_biquad_df1_##S(const T ba[5], int F, const T *clamp, T st[4],           \
                            const T *x, T *y, size_t n) {                            \
        T x1 = st[0], x2 = st[1], y1 = st[2], y2 = st[3];                            \
        for (size_t i = 0; i < n; i++)                                               \
            y[i] = df1c_step_##S(ba, F, clamp, &x1, &x2, &y1, &y2, x[i]);            \
        st[0] = x1; st[1] = x2; st[2] = y1; st[3] = y2;                              \
    }                                                                                \
    void orc_biquad_df1_##S##_lanes(const T ba[5], int F, const T *clamp, T *st,     \
                                    const T *x, T *y, size_t frames, size_t lanes,   \
                                    int layout, int nthreads) {                      \
        LANE_BLOCKS(lanes, nthreads, lo, hi, {                                       \
            T *sx1 = st, *sx2 = st + lanes, *sy1 = st + 2 * lanes,                   \
              *sy2 = st + 3 * lanes;                                                 \
            if (layout == ORC_FRAME_MAJOR) {                                         \
                /* dsp-process/src/compose.rs:468-476 via process.rs:122-127 */      \
                for (size_t t = 0; t < frames; t++) {                                \
                    const T *xr = x + t * lanes;                                     \
                    T *yr = y + t * lanes;                                           \
                    if (clamp) {                                                     \
                        for (size_t l = lo; l < hi; l++)                             \
                            yr[l] = df1c_step_##S(ba, F, clamp, &sx1[l], &sx2[l],    \
                                                  &sy1[l], &sy2[l], xr[l]);          \
                    } else {                                                         \
                        for (size_t l = lo; l < hi; l++)                             \
                            yr[l] = df1_step_##S(ba, F, &sx1[l], &sx2[l], &sy1[l],   \
                                                 &sy2[l], xr[l]);                    \
                    }                                                                \
                }                                                                    \
            } else { /* compose.rs:478-494 */                                        \
                for (size_t l = lo; l < hi; l++) {                                   \
                    T s4[4] = {sx1[l], sx2[l], sy1[l], sy2[l]};                      \
                    orc_biquad_df1_##S(ba, F, clamp, s4, x + l * frames,             \
                                       y + l * frames, frames);                      \
                    sx1[l] = s4[0]; sx2[l] = s4[1]; sy1[l] = s4[2]; sy2[l] = s4[3];  \
                }                                                                    \
            }                                                                        \
        });                                                                          \
    }                                                                                \
    /* biquad.rs:339-364: stage i's y delay line is stage i+1's x delay line */      \
    void orc_biquad_cascade_##S(const T *ba, int F, int nsec, T *st, const T *x,     \
                                T *y, size_t n) {                                    \
        for (size_t i = 0; i < n; i++) {                                             \
            T x0 = x[i];                                                             \
            T *xs = st; /* [x0,x1] of current stage */                               \
            for (int s = 0; s < nsec; s++) {                                         \
                T *ys = st + 2 + 2 * s;                                              \
                T xa = xs[0], xb = xs[1], ya = ys[0], yb = ys[1];                    \
                T t1 = xa, t2 = xb, t3 = ya, t4 = yb;                                \
                T y0 = df1_step_##S(ba + 5 * s, F, &t1, &t2, &t3, &t4, x0);          \
                xs[0] = x0; xs[1] = xa;                                              \
                x0 = y0;                                                             \
                xs = ys;                                                             \
            }                                                                        \
            T prev = xs[0];                                                          \
            xs[0] = x0; xs[1] = prev;                                                \
            y[i] = x0;                                                               \
        }                                                                            \
    }                                                                                \
    void orc_biquad_cascade_##S##_lanes(const T *ba, int F, int nsec, T *st,         \
                                        const T *x, T *y, size_t frames,             \
                                        size_t lanes, int layout, int nthreads) {    \
        int W = 2 + 2 * nsec;                                                        \
        LANE_BLOCKS(lanes, nthreads, lo, hi, {                                       \
            T *ls = (T *)malloc(sizeof(T) * (size_t)W);                              \
            T *xb = (T *)malloc(sizeof(T) * frames);                                 \
            for (size_t l = lo; l < hi; l++) {                                       \
                for (int w = 0; w < W; w++) ls[w] = st[(size_t)w * lanes + l];       \
                if (layout == ORC_FRAME_MAJOR) {                                     \
                    for (size_t t = 0; t < frames; t++) xb[t] = x[t * lanes + l];    \
                    orc_biquad_cascade_##S(ba, F, nsec, ls, xb, xb, frames);         \
                    for (size_t t = 0; t < frames; t++) y[t * lanes + l] = xb[t];    \
                } else {                                                             \
                    orc_biquad_cascade_##S(ba, F, nsec, ls, x + l * frames,          \
                                           y + l * frames, frames);                  \
                }                                                                    \
                for (int w = 0; w < W; w++) st[(size_t)w * lanes + l] = ls[w];       \
            }                                                                        \
            free(ls);                                                                \
            free(xb);                                                                \
        });                                                                          \
    }
DEF_DF1_API(i8, int8_t)
DEF_DF1_API(i16, int16_t)
DEF_DF1_API(i32, int32_t)
DEF_DF1_API(i64, int64_t)
DEF_DF1_API(f32, float)
DEF_DF1_API(f64, double)

/* ------------------------------------------------------------------ */
/* DF2T float: src/iir/biquad.rs:418-440                               */
/* ------------------------------------------------------------------ */
#define DEF_DF2T(S, T)                                                               \
    static inline T df2t_step_##S(const T *ba, const T *cl, T *s0, T *s1, T x0) {    \
        T y0 = *s0 + ba[0] * x0;                                                     \
        if (cl) {                                                                    \
            y0 = y0 + cl[0];                                                         \
            y0 = y0 < cl[1] ? cl[1] : (y0 > cl[2] ? cl[2] : y0);                     \
        }                                                                            \
        *s0 = *s1 + ba[1] * x0 + ba[3] * y0;                                         \
        *s1 = ba[2] * x0 + ba[4] * y0;                                               \
        return y0;                                                                   \
    }                                                                                \
    void orc_biquad_df2t_##S(const T ba[5], const T *clamp, T st[2], const T *x,     \
                             T *y, size_t n) {                                       \
        T s0 = st[0], s1 = st[1];                                                    \
        for (size_t i = 0; i < n; i++) y[i] = df2t_step_##S(ba, clamp, &s0, &s1, x[i]); \
        st[0] = s0; st[1] = s1;                                                      \
    }                                                                                \
    void orc_biquad_df2t_##S##_lanes(const T ba[5], const T *clamp, T *st,           \
                                     const T *x, T *y, size_t frames, size_t lanes,  \
                                     int layout, int nthreads) {                     \
        LANE_BLOCKS(lanes, nthreads, lo, hi, {                                       \
            if (layout == ORC_FRAME_MAJOR) {                                         \
                for (size_t t = 0; t < frames; t++)                                  \
                    for (size_t l = lo; l < hi; l++)                                 \
                        y[t * lanes + l] = df2t_step_##S(ba, clamp, &st[l],          \
                                                         &st[lanes + l],             \
                                                         x[t * lanes + l]);          \
            } else {                                                                 \
                for (size_t l = lo; l < hi; l++) {                                   \
                    T s2[2] = {st[l], st[lanes + l]};                                \
                    orc_biquad_df2t_##S(ba, clamp, s2, x + l * frames,               \
                                        y + l * frames, frames);                     \
                    st[l] = s2[0]; st[lanes + l] = s2[1];                            \
                }                                                                    \
            }                                                                        \
        });                                                                          \
    }
DEF_DF2T(f32, float)
DEF_DF2T(f64, double)

/* ------------------------------------------------------------------ */
/* DirectForm1Wide: src/iir/biquad.rs:445-480 (parity unpinned)        */
/* ------------------------------------------------------------------ */
static inline int32_t wide_step(const int32_t *ba, int F, const int32_t *cl, int32_t *sx,
                                int64_t *sy, int32_t x0) {
    uint64_t acc = (uint64_t)((int64_t)ba[0] * x0) + (uint64_t)((int64_t)ba[1] * sx[0]) +
                   (uint64_t)((int64_t)ba[2] * sx[1]);
    sx[1] = sx[0];
    sx[0] = x0;
    acc += (uint64_t)(((int64_t)(uint64_t)(uint32_t)sy[0] * (int64_t)ba[3]) >> 32);
    acc += (uint64_t)((int64_t)(int32_t)(sy[0] >> 32) * (int64_t)ba[3]);
    acc += (uint64_t)(((int64_t)(uint64_t)(uint32_t)sy[1] * (int64_t)ba[4]) >> 32);
    acc += (uint64_t)((int64_t)(int32_t)(sy[1] >> 32) * (int64_t)ba[4]);
    acc <<= (32 - F);
    sy[1] = sy[0];
    sy[0] = (int64_t)acc;
    int32_t y0 = (int32_t)((int64_t)acc >> 32);
    if (cl) { /* biquad.rs:474-480 */
        int32_t v = (int32_t)((uint32_t)y0 + (uint32_t)cl[0]);
        v = v < cl[1] ? cl[1] : (v > cl[2] ? cl[2] : v);
        sy[0] = (int64_t)(((uint64_t)(int64_t)v << 32) | (uint64_t)(uint32_t)sy[0]);
        return v;
    }
    return y0;
}
static inline void wide_unpack(const int32_t *w, int32_t *sx, int64_t *sy) {
    sx[0] = w[0]; sx[1] = w[1];
    sy[0] = (int64_t)(((uint64_t)(uint32_t)w[3] << 32) | (uint32_t)w[2]);
    sy[1] = (int64_t)(((uint64_t)(uint32_t)w[5] << 32) | (uint32_t)w[4]);
}
static inline void wide_pack(int32_t *w, const int32_t *sx, const int64_t *sy) {
    w[0] = sx[0]; w[1] = sx[1];
    w[2] = (int32_t)(uint32_t)sy[0]; w[3] = (int32_t)(sy[0] >> 32);
    w[4] = (int32_t)(uint32_t)sy[1]; w[5] = (int32_t)(sy[1] >> 32);
}
void orc_biquad_df1wide_i32(const int32_t ba[5], int F, const int32_t *clamp, int32_t st[6],
                            const int32_t *x, int32_t *y, size_t n) {
    int32_t sx[2];
    int64_t sy[2];
    wide_unpack(st, sx, sy);
    for (size_t i = 0; i < n; i++) y[i] = wide_step(ba, F, clamp, sx, sy, x[i]);
    wide_pack(st, sx, sy);
}

/* DirectForm1Dither: src/iir/biquad.rs:484-538 */
static inline int32_t dither_step(const int32_t *ba, int F, const int32_t *cl, int32_t *s,
                                  int32_t x0) {
    uint64_t acc = (uint64_t)(uint32_t)s[4] + (uint64_t)((int64_t)ba[0] * x0) +
                   (uint64_t)((int64_t)ba[1] * s[0]) + (uint64_t)((int64_t)ba[2] * s[1]) +
                   (uint64_t)((int64_t)ba[3] * s[2]) + (uint64_t)((int64_t)ba[4] * s[3]);
    acc <<= (32 - F);
    /* (acc as u32) >> (32-F); for F==0 the low word is 0 after <<32 */
    s[4] = F == 0 ? 0 : (int32_t)(((uint32_t)acc) >> (32 - F));
    int32_t y0 = (int32_t)((int64_t)acc >> 32);
    s[1] = s[0]; s[0] = x0;
    s[3] = s[2]; s[2] = y0;
    if (cl) {
        int32_t v = (int32_t)((uint32_t)y0 + (uint32_t)cl[0]);
        v = v < cl[1] ? cl[1] : (v > cl[2] ? cl[2] : v);
        s[2] = v;
        return v;
    }
    return y0;
}
void orc_biquad_df1dither_i32(const int32_t ba[5], int F, const int32_t *clamp, int32_t st[5],
                              const int32_t *x, int32_t *y, size_t n) {
    for (size_t i = 0; i < n; i++) y[i] = dither_step(ba, F, clamp, st, x[i]);
}

/* generic "gather lane state, run single-lane fn, scatter" lanes driver for i32 word states */
#define DEF_WORDS_LANES(NAME, W, CALL)                                               \
    void NAME##_lanes(const int32_t ba[5], int F, const int32_t *clamp, int32_t *st, \
                      const int32_t *x, int32_t *y, size_t frames, size_t lanes,     \
                      int layout, int nthreads) {                                    \
        LANE_BLOCKS(lanes, nthreads, lo, hi, {                                       \
            int32_t *xb = (int32_t *)malloc(sizeof(int32_t) * (frames ? frames : 1)); \
            for (size_t l = lo; l < hi; l++) {                                       \
                int32_t ls[W];                                                       \
                for (int w = 0; w < W; w++) ls[w] = st[(size_t)w * lanes + l];       \
                if (layout == ORC_FRAME_MAJOR) {                                     \
                    for (size_t t = 0; t < frames; t++) xb[t] = x[t * lanes + l];    \
                    CALL(ba, F, clamp, ls, xb, xb, frames);                          \
                    for (size_t t = 0; t < frames; t++) y[t * lanes + l] = xb[t];    \
                } else {                                                             \
                    CALL(ba, F, clamp, ls, x + l * frames, y + l * frames, frames);  \
                }                                                                    \
                for (int w = 0; w < W; w++) st[(size_t)w * lanes + l] = ls[w];       \
            }                                                                        \
            free(xb);                                                                \
        });                                                                          \
    }
DEF_WORDS_LANES(orc_biquad_df1wide_i32, 6, orc_biquad_df1wide_i32)
DEF_WORDS_LANES(orc_biquad_df1dither_i32, 5, orc_biquad_df1dither_i32)

/* ------------------------------------------------------------------ */
/* src/py.rs:50-108                                                    */
/* ------------------------------------------------------------------ */
void orc_sos(const double *sos, int nsec, int32_t *xy, size_t n) {
    /* [C]::inplace on [S]: stage-major (dsp-process/src/compose.rs:67-77) */
    for (int s = 0; s < nsec; s++) {
        double n5[5];
        int32_t ba[5], st[4] = {0, 0, 0, 0};
        orc_ba_normalize_f64(sos + 6 * s, n5);
        for (int i = 0; i < 5; i++) ba[i] = orc_quant_i32(n5[i], 29);
        orc_biquad_df1_i32(ba, 29, NULL, st, xy, xy, n);
    }
}
void orc_sos_clamp_wide(const double *sos, int nsec, int32_t *xy, size_t n) {
    for (int s = 0; s < nsec; s++) {
        double n5[5];
        int32_t ba[5], cl[3], st[6] = {0, 0, 0, 0, 0, 0};
        orc_ba_normalize_f64(sos + 9 * s, n5);
        for (int i = 0; i < 5; i++) ba[i] = orc_quant_i32(n5[i], 29);
        for (int i = 0; i < 3; i++) cl[i] = orc_round_sat_i32(sos[9 * s + 6 + i]);
        orc_biquad_df1wide_i32(ba, 29, cl, st, xy, xy, n);
    }
}

/* ------------------------------------------------------------------ */
/* Half band filters: src/hbf.rs                                       */
/* ------------------------------------------------------------------ */
/* hbf.rs:308-349 (filter design data of the reference) */
const float ORC_HBF_TAPS0[23] = {
    7.60375795e-07f, -3.77494111e-06f, 1.26458559e-05f, -3.43188253e-05f, 8.10687478e-05f,
    -1.72971467e-04f, 3.40845059e-04f, -6.29522864e-04f, 1.10128831e-03f, -1.83933299e-03f,
    2.95124926e-03f, -4.57290964e-03f, 6.87374176e-03f, -1.00656257e-02f, 1.44199840e-02f,
    -2.03025100e-02f, 2.82462332e-02f, -3.91128509e-02f, 5.44795658e-02f, -7.77002672e-02f,
    1.17523452e-01f, -2.06185388e-01f, 6.34588695e-01f};
const float ORC_HBF_TAPS1[10] = {-1.12811343e-05f, 1.12724671e-04f, -6.07439343e-04f,
                                 2.31904511e-03f, -7.00322950e-03f, 1.78225473e-02f,
                                 -4.01209836e-02f, 8.43315989e-02f, -1.83189521e-01f,
                                 6.26346521e-01f};
const float ORC_HBF_TAPS2[5] = {0.0007686f, -0.00768669f, 0.0386536f, -0.14002434f, 0.60828885f};
const float ORC_HBF_TAPS3[4] = {-0.00261331f, 0.02476858f, -0.12112638f, 0.59897111f};
const float ORC_HBF_TAPS4[3] = {0.01186105f, -0.09808109f, 0.58622005f};

const float *orc_hbf_taps(int idx, int *M) {
    static const float *t[5] = {ORC_HBF_TAPS0, ORC_HBF_TAPS1, ORC_HBF_TAPS2, ORC_HBF_TAPS3,
                                ORC_HBF_TAPS4};
    static const int m[5] = {23, 10, 5, 4, 3};
    if (idx < 0 || idx > 4) return NULL;
    *M = m[idx];
    return t[idx];
}
size_t orc_hbf_dec_state_words(int k) {
    size_t w = 0; int M;
    for (int i = 0; i < k; i++) { orc_hbf_taps(i, &M); w += (size_t)(3 * M - 2); }
    return w;
}
size_t orc_hbf_int_state_words(int k) {
    size_t w = 0; int M;
    for (int i = 0; i < k; i++) { orc_hbf_taps(i, &M); w += (size_t)(2 * M - 1); }
    return w;
}
size_t orc_hbf_dec_response_length(int depth) { /* hbf.rs:424-448 */
    size_t n = 0; int M;
    for (int i = depth - 1; i >= 0; i--) { orc_hbf_taps(i, &M); n /= 2; n += (size_t)(2 * M - 1); }
    return n;
}
size_t orc_hbf_int_response_length(int depth) { /* hbf.rs:515-539 */
    size_t n = 0; int M;
    for (int i = 0; i < depth; i++) { orc_hbf_taps(i, &M); n += (size_t)(2 * M - 1); n *= 2; }
    return n;
}

/* hbf.rs:46-68 for n consecutive windows starting at w[0], w[1], ...: per output the
 * sum runs over taps j = 0..M-1 in order (small taps first, sequential fold, each op
 * rounded).  Iterator::sum::<f32>() folds from -0.0 (Rust >= 1.83; the crate needs
 * >= 1.85), so the result is exactly t0 + t1 + ...  The loops are tap-outer /
 * output-inner so the compiler can vectorise across outputs like LLVM does for the
 * reference; the per-output operation order is unchanged. */
static inline void fir_block(const float *restrict c, int M, int odd, int sym,
                             const float *restrict w, float *restrict out, size_t n) {
    const int top = 2 * M - 1 + odd; /* index of the newest sample of window 0 */
    if (sym) {
        for (size_t i = 0; i < n; i++) out[i] = (w[i + top] + w[i]) * c[0];
        for (int j = 1; j < M; j++) {
            const float cj = c[j];
            for (size_t i = 0; i < n; i++) out[i] = out[i] + (w[i + top - j] + w[i + j]) * cj;
        }
        if (odd)
            for (size_t i = 0; i < n; i++) out[i] = out[i] + w[i + M];
    } else {
        for (size_t i = 0; i < n; i++) out[i] = (w[i + top] - w[i]) * c[0];
        for (int j = 1; j < M; j++) {
            const float cj = c[j];
            for (size_t i = 0; i < n; i++) out[i] = out[i] + (w[i + top - j] - w[i + j]) * cj;
        }
    }
}

#define HBF_CHUNK 64
/* hbf.rs:155-192 restated with explicit history: st = [even(M-1) | odd(2M-1)] */
void orc_hbf_dec_f32(const float *taps, int M, float *st, const float *x, float *y, size_t n) {
    const int LEN = 2 * M - 1;
    float even[ORC_HBF_MAX_M + HBF_CHUNK], odd[2 * ORC_HBF_MAX_M + HBF_CHUNK];
    memcpy(even, st, sizeof(float) * (size_t)(M - 1));
    memcpy(odd, st + (M - 1), sizeof(float) * (size_t)LEN);
    for (size_t o = 0; o < n; o += HBF_CHUNK) {
        size_t c = n - o < HBF_CHUNK ? n - o : HBF_CHUNK;
        for (size_t i = 0; i < c; i++) { /* load input (hbf.rs:167-176) */
            even[M - 1 + i] = x[2 * (o + i)];
            odd[LEN + i] = x[2 * (o + i) + 1];
        }
        float acc[HBF_CHUNK];
        fir_block(taps, M, 0, 1, odd, acc, c);
        for (size_t i = 0; i < c; i++) y[o + i] = acc[i] + even[i]; /* hbf.rs:178-181 */
        memmove(even, even + c, sizeof(float) * (size_t)(M - 1)); /* hbf.rs:183-184 */
        memmove(odd, odd + c, sizeof(float) * (size_t)LEN);
    }
    memcpy(st, even, sizeof(float) * (size_t)(M - 1));
    memcpy(st + (M - 1), odd, sizeof(float) * (size_t)LEN);
}
/* hbf.rs:207-236: st = [x hist (2M-1)] */
void orc_hbf_int_f32(const float *taps, int M, float *st, const float *x, float *y, size_t n) {
    const int LEN = 2 * M - 1;
    float xs[2 * ORC_HBF_MAX_M + HBF_CHUNK];
    memcpy(xs, st, sizeof(float) * (size_t)LEN);
    for (size_t o = 0; o < n; o += HBF_CHUNK) {
        size_t c = n - o < HBF_CHUNK ? n - o : HBF_CHUNK;
        memcpy(xs + LEN, x + o, sizeof(float) * c);
        float acc[HBF_CHUNK];
        fir_block(taps, M, 0, 1, xs, acc, c);
        for (size_t i = 0; i < c; i++) {
            y[2 * (o + i)] = acc[i];
            y[2 * (o + i) + 1] = xs[M + i]; /* center tap: identity */
        }
        memmove(xs, xs + c, sizeof(float) * (size_t)LEN);
    }
    memcpy(st, xs, sizeof(float) * (size_t)LEN);
}
/* hbf.rs:70-138 */
void orc_fir_f32(const float *taps, int M, int odd, int sym, float *st, const float *x, float *y,
                 size_t n) {
    const int LEN = 2 * M - 1 + odd;
    float xs[2 * ORC_HBF_MAX_M + 1 + HBF_CHUNK];
    memcpy(xs, st, sizeof(float) * (size_t)LEN);
    for (size_t o = 0; o < n; o += HBF_CHUNK) {
        size_t c = n - o < HBF_CHUNK ? n - o : HBF_CHUNK;
        memcpy(xs + LEN, x + o, sizeof(float) * c);
        float acc[HBF_CHUNK];
        fir_block(taps, M, odd, sym, xs, acc, c);
        memcpy(y + o, acc, sizeof(float) * c);
        memmove(xs, xs + c, sizeof(float) * (size_t)LEN);
    }
    memcpy(st, xs, sizeof(float) * (size_t)LEN);
}

/* hbf.rs:385-421: /2^k = TAPS[k-1] -> ... -> TAPS[0]; Major scratch chaining is
 * plain sequential composition (dsp-process/src/compose.rs:581-593). Processed
 * here in blocks of 32 output frames like HBF_CASCADE_BLOCK (hbf.rs:357). */
void orc_hbf_dec_cascade_f32(int k, float *st, const float *x, float *y, size_t n_out) {
    enum { B = 32 };
    float a[B * 32], b[B * 16];
    for (size_t o = 0; o < n_out; o += B) {
        size_t c = n_out - o < B ? n_out - o : B;
        const float *src = x + (o << k);
        float *s = st;
        size_t n = c << k; /* samples at current rate */
        float *bufs[2] = {a, b};
        int which = 0;
        for (int i = k - 1; i >= 0; i--) {
            int M;
            const float *t = orc_hbf_taps(i, &M);
            float *dst = (i == 0) ? (y + o) : bufs[which];
            orc_hbf_dec_f32(t, M, s, src, dst, n / 2);
            s += 3 * M - 2;
            src = dst;
            n /= 2;
            which ^= 1;
        }
    }
}
/* hbf.rs:476-512: x2^k = TAPS[0] -> ... -> TAPS[k-1] */
void orc_hbf_int_cascade_f32(int k, float *st, const float *x, float *y, size_t n_in) {
    enum { B = 32 };
    float a[B * 32], b[B * 32];
    for (size_t o = 0; o < n_in; o += B) {
        size_t c = n_in - o < B ? n_in - o : B;
        const float *src = x + o;
        float *s = st;
        size_t n = c;
        float *bufs[2] = {a, b};
        int which = 0;
        for (int i = 0; i < k; i++) {
            int M;
            const float *t = orc_hbf_taps(i, &M);
            float *dst = (i == k - 1) ? (y + (o << k)) : bufs[which];
            orc_hbf_int_f32(t, M, s, src, dst, n);
            s += 2 * M - 1;
            src = dst;
            n *= 2;
            which ^= 1;
        }
    }
}

void orc_hbf_dec_cascade_f32_lanes(int k, float *st, const float *x, float *y, size_t n_out,
                                   size_t lanes, int layout, int nthreads) {
    const size_t W = orc_hbf_dec_state_words(k), R = (size_t)1 << k;
    LANE_BLOCKS(lanes, nthreads, lo, hi, {
        float *ls = (float *)malloc(sizeof(float) * W);
        float xb[32 * 32], yb[32];
        for (size_t l = lo; l < hi; l++) {
            for (size_t w = 0; w < W; w++) ls[w] = st[w * lanes + l];
            if (layout == ORC_FRAME_MAJOR) {
                for (size_t o = 0; o < n_out; o += 32) {
                    size_t c = n_out - o < 32 ? n_out - o : 32;
                    for (size_t t = 0; t < c; t++)
                        memcpy(xb + t * R, x + ((o + t) * lanes + l) * R, sizeof(float) * R);
                    orc_hbf_dec_cascade_f32(k, ls, xb, yb, c);
                    for (size_t t = 0; t < c; t++) y[(o + t) * lanes + l] = yb[t];
                }
            } else {
                orc_hbf_dec_cascade_f32(k, ls, x + l * n_out * R, y + l * n_out, n_out);
            }
            for (size_t w = 0; w < W; w++) st[w * lanes + l] = ls[w];
        }
        free(ls);
    });
}
void orc_hbf_int_cascade_f32_lanes(int k, float *st, const float *x, float *y, size_t n_in,
                                   size_t lanes, int layout, int nthreads) {
    const size_t W = orc_hbf_int_state_words(k), R = (size_t)1 << k;
    LANE_BLOCKS(lanes, nthreads, lo, hi, {
        float *ls = (float *)malloc(sizeof(float) * W);
        float xb[32], yb[32 * 32];
        for (size_t l = lo; l < hi; l++) {
            for (size_t w = 0; w < W; w++) ls[w] = st[w * lanes + l];
            if (layout == ORC_FRAME_MAJOR) {
                for (size_t o = 0; o < n_in; o += 32) {
                    size_t c = n_in - o < 32 ? n_in - o : 32;
                    for (size_t t = 0; t < c; t++) xb[t] = x[(o + t) * lanes + l];
                    orc_hbf_int_cascade_f32(k, ls, xb, yb, c);
                    for (size_t t = 0; t < c; t++)
                        memcpy(y + ((o + t) * lanes + l) * R, yb + t * R, sizeof(float) * R);
                }
            } else {
                orc_hbf_int_cascade_f32(k, ls, x + l * n_in, y + l * n_in * R, n_in);
            }
            for (size_t w = 0; w < W; w++) st[w * lanes + l] = ls[w];
        }
        free(ls);
    });
}

/* ------------------------------------------------------------------ */
/* Lowpass: src/lowpass.rs:47-78 (parity unpinned in the reference)     */
/* ------------------------------------------------------------------ */
static inline int32_t lowpass_step(int order, const int32_t *k, int64_t *s, int32_t x) {
    uint64_t d = (uint64_t)((int64_t)sat_sub_i32(x, (int32_t)(s[0] >> 32)) * (int64_t)k[0]);
    int32_t y;
    if (order == 1) {
        s[0] = (int64_t)((uint64_t)s[0] + d);
        y = (int32_t)(s[0] >> 32);
        s[0] = (int64_t)((uint64_t)s[0] + d);
    } else {
        d += (uint64_t)(s[1] >> 32) * (uint64_t)(int64_t)k[1];
        s[1] = (int64_t)((uint64_t)s[1] + d);
        s[0] = (int64_t)((uint64_t)s[0] + (uint64_t)s[1]);
        y = (int32_t)(s[0] >> 32);
        s[0] = (int64_t)((uint64_t)s[0] + (uint64_t)s[1]);
        s[1] = (int64_t)((uint64_t)s[1] + d);
    }
    return y;
}
void orc_lowpass_i32(int order, const int32_t *k, int64_t *st, const int32_t *x, int32_t *y,
                     size_t n) {
    for (size_t i = 0; i < n; i++) y[i] = lowpass_step(order, k, st, x[i]);
}
void orc_lowpass_i32_lanes(int order, const int32_t *k, int64_t *st, const int32_t *x, int32_t *y,
                           size_t frames, size_t lanes, int layout, int nthreads) {
    LANE_BLOCKS(lanes, nthreads, lo, hi, {
        for (size_t l = lo; l < hi; l++) {
            int64_t s[2] = {st[l], order == 2 ? st[lanes + l] : 0};
            for (size_t t = 0; t < frames; t++) {
                size_t i = layout == ORC_FRAME_MAJOR ? t * lanes + l : l * frames + t;
                y[i] = lowpass_step(order, k, s, x[i]);
            }
            st[l] = s[0];
            if (order == 2) st[lanes + l] = s[1];
        }
    });
}

/* Accu (src/accu.rs:34-37) -> Complex::from_angle (src/complex.rs:237-240) ->
 * Lockin (src/lockin.rs:17-39), mix = i32 * Q32<32> (dsp-fixedpoint/src/lib.rs:449-456) */
void orc_lockin_i32_lanes(int order, const int32_t *k, int32_t *accu_state,
                          const int32_t *accu_step, int64_t *lp_st, const int32_t *x, int32_t *iq,
                          size_t frames, size_t lanes, int layout, int nthreads) {
    make_tables();
    LANE_BLOCKS(lanes, nthreads, lo, hi, {
        for (size_t l = lo; l < hi; l++) {
            int64_t si[2] = {0, 0}, sq[2] = {0, 0};
            for (int w = 0; w < order; w++) {
                si[w] = lp_st[(size_t)w * lanes + l];
                sq[w] = lp_st[(size_t)(order + w) * lanes + l];
            }
            uint32_t ph = (uint32_t)accu_state[l], step = (uint32_t)accu_step[l];
            for (size_t t = 0; t < frames; t++) {
                size_t i = layout == ORC_FRAME_MAJOR ? t * lanes + l : l * frames + t;
                ph += step;
                int32_t c, s;
                cossin_tab(g_cossin, (int32_t)ph, &c, &s);
                int32_t mi = (int32_t)(((int64_t)c * (int64_t)x[i]) >> 32);
                int32_t mq = (int32_t)(((int64_t)s * (int64_t)x[i]) >> 32);
                iq[2 * i] = lowpass_step(order, k, si, mi);
                iq[2 * i + 1] = lowpass_step(order, k, sq, mq);
            }
            accu_state[l] = (int32_t)ph;
            for (int w = 0; w < order; w++) {
                lp_st[(size_t)w * lanes + l] = si[w];
                lp_st[(size_t)(order + w) * lanes + l] = sq[w];
            }
        }
    });
}

/* Lockin on (sample, phase) tuples: src/lockin.rs:30-39 (phase -> Complex::from_angle -> Q32<32> LO, then
 * the (X, Complex<U>) impl).  xp = (x, phase) pairs, pair innermost. */
void orc_lockin_phase_i32_lanes(int order, const int32_t *k, int64_t *lp_st, const int32_t *xp, int32_t *iq,
                                size_t frames, size_t lanes, int layout, int nthreads) {
    make_tables();
    LANE_BLOCKS(lanes, nthreads, lo, hi, {
        for (size_t l = lo; l < hi; l++) {
            int64_t si[2] = {0, 0}, sq[2] = {0, 0};
            for (int w = 0; w < order; w++) {
                si[w] = lp_st[(size_t)w * lanes + l];
                sq[w] = lp_st[(size_t)(order + w) * lanes + l];
            }
            for (size_t t = 0; t < frames; t++) {
                size_t i = layout == ORC_FRAME_MAJOR ? t * lanes + l : l * frames + t;
                int32_t c, s;
                cossin_tab(g_cossin, xp[2 * i + 1], &c, &s);
                int32_t mi = (int32_t)(((int64_t)c * (int64_t)xp[2 * i]) >> 32);
                int32_t mq = (int32_t)(((int64_t)s * (int64_t)xp[2 * i]) >> 32);
                iq[2 * i] = lowpass_step(order, k, si, mi);
                iq[2 * i + 1] = lowpass_step(order, k, sq, mq);
            }
            for (int w = 0; w < order; w++) {
                lp_st[(size_t)w * lanes + l] = si[w];
                lp_st[(size_t)(order + w) * lanes + l] = sq[w];
            }
        }
    });
}

/* Lockin on (sample, LO) tuples: src/lockin.rs:17-28 with X = i32, U = Q32<32>; `x * lo.re()` is
 * i32 * Q32<32> = ((x as i64 * lo as i64) >> 32) as i32 (dsp-fixedpoint/src/lib.rs:449-456).
 * xlo = (x, lo.re, lo.im) triples. */
void orc_lockin_lo_i32_lanes(int order, const int32_t *k, int64_t *lp_st, const int32_t *xlo, int32_t *iq,
                             size_t frames, size_t lanes, int layout, int nthreads) {
    LANE_BLOCKS(lanes, nthreads, lo, hi, {
        for (size_t l = lo; l < hi; l++) {
            int64_t si[2] = {0, 0}, sq[2] = {0, 0};
            for (int w = 0; w < order; w++) {
                si[w] = lp_st[(size_t)w * lanes + l];
                sq[w] = lp_st[(size_t)(order + w) * lanes + l];
            }
            for (size_t t = 0; t < frames; t++) {
                size_t i = layout == ORC_FRAME_MAJOR ? t * lanes + l : l * frames + t;
                int32_t mi = (int32_t)(((int64_t)xlo[3 * i + 1] * (int64_t)xlo[3 * i]) >> 32);
                int32_t mq = (int32_t)(((int64_t)xlo[3 * i + 2] * (int64_t)xlo[3 * i]) >> 32);
                iq[2 * i] = lowpass_step(order, k, si, mi);
                iq[2 * i + 1] = lowpass_step(order, k, sq, mq);
            }
            for (int w = 0; w < order; w++) {
                lp_st[(size_t)w * lanes + l] = si[w];
                lp_st[(size_t)(order + w) * lanes + l] = sq[w];
            }
        }
    });
}

/* config 5 chain: HbfDec(/2^k) -> HbfInt(x2^k) -> Biquad<f32> DF1 */
void orc_chain_f32_lanes(int k, const float ba[5], float *st, const float *x, float *y,
                         size_t n_low, size_t lanes, int layout, int nthreads) {
    const size_t WD = orc_hbf_dec_state_words(k), WI = orc_hbf_int_state_words(k);
    const size_t W = WD + WI + 4, R = (size_t)1 << k;
    LANE_BLOCKS(lanes, nthreads, lo, hi, {
        float *ls = (float *)malloc(sizeof(float) * W);
        float xb[32 * 32], lowb[32], yb[32 * 32];
        for (size_t l = lo; l < hi; l++) {
            for (size_t w = 0; w < W; w++) ls[w] = st[w * lanes + l];
            for (size_t o = 0; o < n_low; o += 32) {
                size_t c = n_low - o < 32 ? n_low - o : 32;
                const float *src;
                if (layout == ORC_FRAME_MAJOR) {
                    for (size_t t = 0; t < c; t++)
                        memcpy(xb + t * R, x + ((o + t) * lanes + l) * R, sizeof(float) * R);
                    src = xb;
                } else {
                    src = x + (l * n_low + o) * R;
                }
                orc_hbf_dec_cascade_f32(k, ls, src, lowb, c);
                orc_hbf_int_cascade_f32(k, ls + WD, lowb, yb, c);
                orc_biquad_df1_f32(ba, 0, NULL, ls + WD + WI, yb, yb, c * R);
                if (layout == ORC_FRAME_MAJOR) {
                    for (size_t t = 0; t < c; t++)
                        memcpy(y + ((o + t) * lanes + l) * R, yb + t * R, sizeof(float) * R);
                } else {
                    memcpy(y + (l * n_low + o) * R, yb, sizeof(float) * c * R);
                }
            }
            for (size_t w = 0; w < W; w++) st[w * lanes + l] = ls[w];
        }
        free(ls);
    });
}

/* ------------------------------------------------------------------ */
/* Cic<T, N, M> (src/cic.rs:13-200), SURVEY 8(f) rank 3.                */
/* State words per lane (type T), ABI order:                            */
/*   [0] index  [1] zoh  [2 + n*M + m] combs[n][m]  [2 + N*M + n] integrators[n]  */
/* Integer adds wrap (decimator: wrapping_add/sub, cic.rs:183-197; the  */
/* interpolator's `+=`/`-` wrap in release builds, cic.rs:156-170).     */
/* ------------------------------------------------------------------ */
#define ORC_CIC_MAXN 8
#define ORC_CIC_MAXM 8
size_t orc_cic_state_words(int N, int M) { return (size_t)(2 + N * M + N); }

#define DEF_CIC(SUF, T, UT)                                                                         \
    typedef struct {                                                                                \
        uint32_t rate, index;                                                                       \
        T zoh, combs[ORC_CIC_MAXN][ORC_CIC_MAXM], integ[ORC_CIC_MAXN];                              \
    } cic_##SUF;                                                                                    \
    static void cic_load_##SUF(cic_##SUF *c, int N, int M, uint32_t rate, const T *st, size_t stride, size_t l) { \
        c->rate = rate;                                                                             \
        c->index = (uint32_t)st[l];                                                                 \
        c->zoh = st[stride + l];                                                                    \
        for (int n = 0; n < N; n++)                                                                 \
            for (int m = 0; m < M; m++) c->combs[n][m] = st[(size_t)(2 + n * M + m) * stride + l];  \
        for (int n = 0; n < N; n++) c->integ[n] = st[(size_t)(2 + N * M + n) * stride + l];         \
    }                                                                                               \
    static void cic_store_##SUF(const cic_##SUF *c, int N, int M, T *st, size_t stride, size_t l) { \
        st[l] = (T)c->index;                                                                        \
        st[stride + l] = c->zoh;                                                                    \
        for (int n = 0; n < N; n++)                                                                 \
            for (int m = 0; m < M; m++) st[(size_t)(2 + n * M + m) * stride + l] = c->combs[n][m];  \
        for (int n = 0; n < N; n++) st[(size_t)(2 + N * M + n) * stride + l] = c->integ[n];         \
    }                                                                                               \
    /* comb cascade, cic.rs:159-164 / :190-196 */                                                   \
    static T cic_combs_##SUF(cic_##SUF *c, int N, int M, T x) {                                     \
        for (int n = 0; n < N; n++) {                                                               \
            T y = (T)((UT)x - (UT)c->combs[n][0]);                                                  \
            for (int m = 0; m + 1 < M; m++) c->combs[n][m] = c->combs[n][m + 1];                    \
            c->combs[n][M - 1] = x;                                                                 \
            x = y;                                                                                  \
        }                                                                                           \
        return x;                                                                                   \
    }                                                                                               \
    /* Process<T, Option<T>> (decimator), cic.rs:176-200; returns 1 and *y on a tick */             \
    static int cic_dec_step_##SUF(cic_##SUF *c, int N, int M, T x, T *y) {                          \
        for (int n = 0; n < N; n++) {                                                               \
            c->integ[n] = (T)((UT)c->integ[n] + (UT)x);                                             \
            x = c->integ[n];                                                                        \
        }                                                                                           \
        if (c->index > 0) {                                                                         \
            c->index -= 1;                                                                          \
            return 0;                                                                               \
        }                                                                                           \
        c->index = c->rate;                                                                         \
        c->zoh = cic_combs_##SUF(c, N, M, x);                                                       \
        *y = c->zoh;                                                                                \
        return 1;                                                                                   \
    }                                                                                               \
    /* Process<Option<T>, T> (interpolator), cic.rs:149-172 */                                      \
    static T cic_int_step_##SUF(cic_##SUF *c, int N, int M, int some, T x) {                        \
        if (some) {                                                                                 \
            c->index = c->rate;                                                                     \
            c->zoh = cic_combs_##SUF(c, N, M, x);                                                   \
        } else {                                                                                    \
            c->index -= 1;                                                                          \
        }                                                                                           \
        T v = c->zoh;                                                                               \
        for (int n = 0; n < N; n++) {                                                               \
            c->integ[n] = (T)((UT)c->integ[n] + (UT)v);                                             \
            v = c->integ[n];                                                                        \
        }                                                                                           \
        return v;                                                                                   \
    }                                                                                               \
    /* Decimator adapter (dsp-process/src/adapters.rs:154-222): frames of R = rate+1 inputs,        \
     * one output per frame = the value of the single tick (zoh) */                                 \
    void orc_cic_dec_##SUF##_lanes(int N, int M, uint32_t rate, T *st, const T *x, T *y, size_t frames, \
                                   size_t lanes, int layout, int nthreads) {                        \
        const size_t R = (size_t)rate + 1;                                                          \
        LANE_BLOCKS(lanes, nthreads, lo, hi, {                                                      \
            for (size_t l = lo; l < hi; l++) {                                                      \
                cic_##SUF c;                                                                        \
                cic_load_##SUF(&c, N, M, rate, st, lanes, l);                                       \
                for (size_t t = 0; t < frames; t++) {                                               \
                    size_t f = layout == ORC_FRAME_MAJOR ? t * lanes + l : l * frames + t;          \
                    T out = c.zoh;                                                                  \
                    for (size_t j = 0; j < R; j++) {                                                \
                        T v;                                                                        \
                        if (cic_dec_step_##SUF(&c, N, M, x[f * R + j], &v)) out = v;                \
                    }                                                                               \
                    y[f] = out;                                                                     \
                }                                                                                   \
                cic_store_##SUF(&c, N, M, st, lanes, l);                                            \
            }                                                                                       \
        });                                                                                         \
    }                                                                                               \
    /* Interpolator adapter (adapters.rs:27-35): Some(x) then R-1 times None */                     \
    void orc_cic_int_##SUF##_lanes(int N, int M, uint32_t rate, T *st, const T *x, T *y, size_t frames, \
                                   size_t lanes, int layout, int nthreads) {                        \
        const size_t R = (size_t)rate + 1;                                                          \
        LANE_BLOCKS(lanes, nthreads, lo, hi, {                                                      \
            for (size_t l = lo; l < hi; l++) {                                                      \
                cic_##SUF c;                                                                        \
                cic_load_##SUF(&c, N, M, rate, st, lanes, l);                                       \
                for (size_t t = 0; t < frames; t++) {                                               \
                    size_t f = layout == ORC_FRAME_MAJOR ? t * lanes + l : l * frames + t;          \
                    for (size_t j = 0; j < R; j++) y[f * R + j] = cic_int_step_##SUF(&c, N, M, j == 0, x[f]); \
                }                                                                                   \
                cic_store_##SUF(&c, N, M, st, lanes, l);                                            \
            }                                                                                       \
        });                                                                                         \
    }
DEF_CIC(i32, int32_t, uint32_t)
DEF_CIC(i64, int64_t, uint64_t)

/* gain() = (M*(rate+1))^N (cic.rs:99-101), gain_log2() (cic.rs:107-109), response_length() (:112-114) */
int64_t orc_cic_gain(int N, int M, uint32_t rate) {
    uint64_t g = 1, b = (uint64_t)M * ((uint64_t)rate + 1);
    for (int n = 0; n < N; n++) g *= b;
    return (int64_t)g;
}
uint32_t orc_cic_gain_log2(int N, int M, uint32_t rate) {
    uint32_t v = (uint32_t)M * rate + (uint32_t)(M - 1);
    uint32_t lz = v ? (uint32_t)__builtin_clz(v) : 32u;
    return (32u - lz) * (uint32_t)N;
}
size_t orc_cic_response_length(int N, uint32_t rate) { return (size_t)rate * (size_t)N; }

/* ------------------------------------------------------------------ */
/* PLL (src/pll.rs:33-108), SURVEY 8(f) rank 4.  All math is wrapping   */
/* 32/64-bit integer.  State words per lane (i32), ABI order:           */
/*   [0] clamp.x0  [1] clamp.clamp (-1,0,1)  [2] z0  [3] y0             */
/*   [4] f0 lo  [5] f0 hi  [6] f lo  [7] f hi  [8] y                    */
/* ------------------------------------------------------------------ */
/* f32 -> Q32<32>: (v * 2^32).round() as i32, f32 arithmetic, saturating cast
 * (dsp-fixedpoint/src/num_traits_impl.rs:30-45) */
static int32_t q32_32_from_f32(float v) {
    float s = roundf(v * 4294967296.0f);
    if (s != s) return 0;
    if (s >= 2147483648.0f) return INT32_MAX;
    if (s <= -2147483648.0f) return INT32_MIN;
    return (int32_t)s;
}
/* PLL::from_bandwidth / from_zpk (src/pll.rs:41-57), f32 arithmetic like the reference */
void orc_pll_from_bandwidth(float bw, float split, int32_t ba[3]) {
    float a = bw * 2.0f * 3.14159274101257324f; /* core::f32::consts::PI */
    float z = 1.0f - a / split;
    float p = 1.0f - a * split;
    float k = -a * a * split;
    ba[0] = q32_32_from_f32(k);
    ba[1] = q32_32_from_f32(-k * z);
    ba[2] = q32_32_from_f32(-(1.0f - p));
}
typedef struct {
    int32_t x0, clamp, z0, y0;
    int64_t f0, f;
    int32_t y;
} pll_state;
/* SplitProcess<W<i32>, W<i32>, PLLState> for PLL (src/pll.rs:88-108) with ClampWrap
 * (src/unwrap.rs:166-194) and overflowing_sub (src/unwrap.rs:73-81) */
static int32_t pll_step(const int32_t ba[3], pll_state *s, int32_t x) {
    s->y = (int32_t)((uint32_t)s->y + (uint32_t)(int32_t)(s->f >> 32));
    int32_t t = (int32_t)((uint32_t)x + (uint32_t)s->y);
    int32_t delta = (int32_t)((uint32_t)t - (uint32_t)s->x0);
    int a = delta >= 0, b = t >= s->x0;
    int wrap = a > b ? 1 : (a < b ? -1 : 0);
    s->x0 = t;
    int c = s->clamp + wrap;
    s->clamp = c > 0 ? 1 : (c < 0 ? -1 : 0);
    int32_t o = s->clamp < 0 ? INT32_MIN : (s->clamp > 0 ? INT32_MAX : t);
    int32_t z0 = o >> 1;
    int32_t y0 = (int32_t)((uint32_t)z0 + (uint32_t)s->z0);
    s->z0 = z0;
    uint64_t acc = (uint64_t)((int64_t)ba[0] * y0) + (uint64_t)((int64_t)ba[1] * s->y0) +
                   (uint64_t)((int64_t)ba[2] * (int32_t)(s->f0 >> 32));
    acc += (uint64_t)(((int64_t)ba[2] * (int64_t)(uint32_t)s->f0) >> 32);
    s->f0 = (int64_t)((uint64_t)s->f0 + acc);
    s->y0 = y0;
    s->f = (int64_t)((uint64_t)s->f + (uint64_t)s->f0);
    return s->y;
}
void orc_pll_i32_lanes(const int32_t ba[3], int32_t *st, const int32_t *x, int32_t *y, size_t frames,
                       size_t lanes, int layout, int nthreads) {
    LANE_BLOCKS(lanes, nthreads, lo, hi, {
        for (size_t l = lo; l < hi; l++) {
            pll_state s;
            s.x0 = st[l]; s.clamp = st[lanes + l]; s.z0 = st[2 * lanes + l]; s.y0 = st[3 * lanes + l];
            s.f0 = (int64_t)((uint64_t)(uint32_t)st[4 * lanes + l] | ((uint64_t)(uint32_t)st[5 * lanes + l] << 32));
            s.f = (int64_t)((uint64_t)(uint32_t)st[6 * lanes + l] | ((uint64_t)(uint32_t)st[7 * lanes + l] << 32));
            s.y = st[8 * lanes + l];
            for (size_t t = 0; t < frames; t++) {
                size_t i = layout == ORC_FRAME_MAJOR ? t * lanes + l : l * frames + t;
                y[i] = pll_step(ba, &s, x[i]);
            }
            st[l] = s.x0; st[lanes + l] = s.clamp; st[2 * lanes + l] = s.z0; st[3 * lanes + l] = s.y0;
            st[4 * lanes + l] = (int32_t)(uint32_t)s.f0; st[5 * lanes + l] = (int32_t)(s.f0 >> 32);
            st[6 * lanes + l] = (int32_t)(uint32_t)s.f; st[7 * lanes + l] = (int32_t)(s.f >> 32);
            st[8 * lanes + l] = s.y;
        }
    });
}

/* ------------------------------------------------------------------ */
/* FM discriminator graph, examples/fm_disc.rs:26-48: */
/* z = x * prev.into_bits().conj() (src/complex.rs:117-134: wide products, */
/* `.as_()` = >> 32), d = atan2(z.im, z.re) - carrier (src/complex.rs:254),*/
/* y = Biquad<Q32<F>> DF1 (src/iir/biquad.rs:366-383).                     */
/* State words (i32): [has_prev, prev.re, prev.im, x1, x2, y1, y2].        */
/* ------------------------------------------------------------------ */
void orc_fm_disc_i32_lanes(int32_t carrier, const int32_t ba[5], int F, int32_t *st, const int32_t *x,
                           int32_t *y, size_t frames, size_t lanes, int layout, int nthreads) {
    make_tables();
    LANE_BLOCKS(lanes, nthreads, lo, hi, {
        for (size_t l = lo; l < hi; l++) {
            int32_t has = st[l], pre = st[lanes + l], pim = st[2 * lanes + l];
            int32_t s[4] = {st[3 * lanes + l], st[4 * lanes + l], st[5 * lanes + l], st[6 * lanes + l]};
            for (size_t t = 0; t < frames; t++) {
                size_t i = layout == ORC_FRAME_MAJOR ? t * lanes + l : l * frames + t;
                int32_t xre = x[2 * i], xim = x[2 * i + 1], d = 0;
                if (has) {
                    int32_t cim = (int32_t)(0u - (uint32_t)pim);
                    int64_t re = (int64_t)((uint64_t)((int64_t)xre * pre) - (uint64_t)((int64_t)xim * cim));
                    int64_t im = (int64_t)((uint64_t)((int64_t)xre * cim) + (uint64_t)((int64_t)xim * pre));
                    d = (int32_t)((uint32_t)orc_atan2((int32_t)(im >> 32), (int32_t)(re >> 32)) - (uint32_t)carrier);
                }
                has = 1; pre = xre; pim = xim;
                orc_biquad_df1_i32(ba, F, NULL, s, &d, &y[i], 1);
            }
            st[l] = has; st[lanes + l] = pre; st[2 * lanes + l] = pim;
            st[3 * lanes + l] = s[0]; st[4 * lanes + l] = s[1]; st[5 * lanes + l] = s[2]; st[6 * lanes + l] = s[3];
        }
    });
}
