// coeff.cu -- host-side coefficient builders of the C ABI (include/idsp_b200.h).
// No device code: restates iir::coefficients::Filter (src/iir/coefficients.rs:111-527), the Biquad
// `From` conversions / from_zpk (src/iir/biquad.rs:545-619), the float -> Q quantisation
// (dsp-fixedpoint/src/num_traits_impl.rs:32-45) and pid::Builder::build (src/iir/pid.rs:193-303) in the
// reference's two float widths.  Every arithmetic step is written in the reference's order so that the f32
// flavour rounds where `Filter<f32>` rounds (host code: g++ on x86-64 does not contract a*b+c).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <limits>
#include <type_traits>

#include "common.cuh"

namespace {

template <class T> struct Fm;  // libm in the width of T (Rust std: f32::sin -> sinf, ...)
template <> struct Fm<double> {
    static double sin(double v) { return ::sin(v); }
    static double cos(double v) { return ::cos(v); }
    static double sinh(double v) { return ::sinh(v); }
    static double sqrt(double v) { return ::sqrt(v); }
    static double round(double v) { return ::round(v); }
    static constexpr double LN_2 = 0.693147180559945309417232121458176568;
    static constexpr double PI = 3.14159265358979323846264338327950288;
    static constexpr double SQRT_2 = 1.41421356237309504880168872420969808;
};
template <> struct Fm<float> {
    static float sin(float v) { return ::sinf(v); }
    static float cos(float v) { return ::cosf(v); }
    static float sinh(float v) { return ::sinhf(v); }
    static float sqrt(float v) { return ::sqrtf(v); }
    static float round(float v) { return ::roundf(v); }
    static constexpr float LN_2 = 0.693147180559945309417232121458176568f;
    static constexpr float PI = 3.14159265358979323846264338327950288f;
    static constexpr float SQRT_2 = 1.41421356237309504880168872420969808f;
};

// Rust `as`: float -> int saturates, NaN -> 0
template <class I, class T> I sat_cast(T v) {
    if (v != v) return 0;
    // 2^(bits-1) is exactly representable in float and double
    const T hi = (T)ldexp(1.0, (int)(8 * sizeof(I) - 1));
    if (v >= hi) return std::numeric_limits<I>::max();
    if (v < -hi) return std::numeric_limits<I>::min();
    return (I)v;  // in range: truncation of an already integral value
}

// `AsPrimitive<Q<T,A,F>> for f`: (v * 2^F).round() as T
template <class I, class T> I quantize(T v, int F) {
    const T scale = (T)ldexp(1.0, F);  // `const { 1.0 / DELTA as $ty }`: exact power of two in T
    return sat_cast<I, T>(Fm<T>::round(v * scale));
}

// one value of float type T -> coefficient type `kind`, stored at out[i]
template <class T> int put(void *out, int i, int kind, int F, T v) {
    switch (kind) {
        case IDSP_I8: ((int8_t *)out)[i] = quantize<int8_t, T>(v, F); return 0;
        case IDSP_I16: ((int16_t *)out)[i] = quantize<int16_t, T>(v, F); return 0;
        case IDSP_I32: ((int32_t *)out)[i] = quantize<int32_t, T>(v, F); return 0;
        case IDSP_I64: ((int64_t *)out)[i] = quantize<int64_t, T>(v, F); return 0;
        case IDSP_F32: ((float *)out)[i] = (float)v; return 0;
        case IDSP_F64: ((double *)out)[i] = (double)v; return 0;
    }
    return -1;
}

bool kind_ok(int kind, int F) {
    if (kind < IDSP_I8 || kind > IDSP_F64) return false;
    if (kind <= IDSP_I64) return F >= -128 && F <= 127;  // const F: i8
    return true;
}

template <class T, class FT> struct FilterImpl {
    // coefficients.rs:241-265
    static int validate(const FT *f) {
        auto fin = [](T v) { return std::isfinite(v); };
        if (!fin(f->frequency)) return idsp_set_error("NonFinite(frequency)"), IDSP_EINVAL;
        if (f->frequency < (T)0 || f->frequency > Fm<T>::PI) return idsp_set_error("OutOfRange(frequency)"), IDSP_EINVAL;
        if (!fin(f->gain) || f->gain <= (T)0) return idsp_set_error("NonPositive(gain)"), IDSP_EINVAL;
        if (!fin(f->shelf) || f->shelf <= (T)0) return idsp_set_error("NonPositive(shelf)"), IDSP_EINVAL;
        switch (f->shape_kind) {
            case IDSP_SHAPE_Q:
                if (!fin(f->shape)) return idsp_set_error("NonFinite(q)"), IDSP_EINVAL;
                if (f->shape <= (T)0) return idsp_set_error("NonPositive(q)"), IDSP_EINVAL;
                return IDSP_OK;
            case IDSP_SHAPE_BANDWIDTH:
                if (!fin(f->shape)) return idsp_set_error("NonFinite(bandwidth)"), IDSP_EINVAL;
                return IDSP_OK;
            case IDSP_SHAPE_SLOPE:
                if (!fin(f->shape)) return idsp_set_error("NonFinite(slope)"), IDSP_EINVAL;
                if (f->shape <= (T)0) return idsp_set_error("NonPositive(slope)"), IDSP_EINVAL;
                return IDSP_OK;
        }
        return idsp_set_error("shape_kind must be 0 (Q), 1 (Bandwidth) or 2 (Slope)"), IDSP_EINVAL;
    }
    // coefficients.rs:266-283
    static T qi(const FT *f) {
        switch (f->shape_kind) {
            case IDSP_SHAPE_BANDWIDTH:
                return (T)2.0 * Fm<T>::sinh(Fm<T>::LN_2 / (T)2.0 * f->shape * f->frequency / Fm<T>::sin(f->frequency));
            case IDSP_SHAPE_SLOPE:
                return Fm<T>::sqrt((f->shelf + (T)1 / f->shelf) * ((T)1 / f->shape - (T)1) + (T)2.0);
            default:
                return (T)1 / f->shape;
        }
    }
    // coefficients.rs:302-479; out = [b0 b1 b2 a0 a1 a2]
    static int build(const FT *f, int type, T *o) {
        const T fsin = Fm<T>::sin(f->frequency), fcos = Fm<T>::cos(f->frequency);
        const T alpha = (T)0.5 * fsin * qi(f);
        const T g = f->gain, one = (T)1, two = (T)2.0, m2 = (T)-2.0;
        switch (type) {
            case IDSP_LOWPASS: {
                const T b = g * (T)0.5 * (one - fcos);
                o[0] = b; o[1] = two * b; o[2] = b;
                o[3] = one + alpha; o[4] = m2 * fcos; o[5] = one - alpha;
                return IDSP_OK;
            }
            case IDSP_HIGHPASS: {
                const T b = g * (T)0.5 * (one + fcos);
                o[0] = b; o[1] = m2 * b; o[2] = b;
                o[3] = one + alpha; o[4] = m2 * fcos; o[5] = one - alpha;
                return IDSP_OK;
            }
            case IDSP_BANDPASS: {
                const T b = g * alpha;
                o[0] = b; o[1] = (T)0; o[2] = -b;
                o[3] = one + alpha; o[4] = m2 * fcos; o[5] = one - alpha;
                return IDSP_OK;
            }
            case IDSP_NOTCH: {
                const T f2 = m2 * fcos;
                o[0] = g; o[1] = f2 * g; o[2] = g;
                o[3] = one + alpha; o[4] = f2; o[5] = one - alpha;
                return IDSP_OK;
            }
            case IDSP_ALLPASS: {
                const T f2 = m2 * fcos;
                o[0] = (one - alpha) * g; o[1] = f2 * g; o[2] = (one + alpha) * g;
                o[3] = one + alpha; o[4] = f2; o[5] = one - alpha;
                return IDSP_OK;
            }
            case IDSP_PEAKING: {
                const T s = Fm<T>::sqrt(f->shelf), f2 = m2 * fcos;
                o[0] = (one + alpha * s) * g; o[1] = f2 * g; o[2] = (one - alpha * s) * g;
                o[3] = one + alpha / s; o[4] = f2; o[5] = one - alpha / s;
                return IDSP_OK;
            }
            case IDSP_LOWSHELF: {
                const T s = Fm<T>::sqrt(f->shelf);
                const T tsa = two * Fm<T>::sqrt(s) * alpha, sp1 = s + one, sm1 = s - one;
                o[0] = s * g * (sp1 - sm1 * fcos + tsa);
                o[1] = two * s * g * (sm1 - sp1 * fcos);
                o[2] = s * g * (sp1 - sm1 * fcos - tsa);
                o[3] = sp1 + sm1 * fcos + tsa;
                o[4] = m2 * (sm1 + sp1 * fcos);
                o[5] = sp1 + sm1 * fcos - tsa;
                return IDSP_OK;
            }
            case IDSP_HIGHSHELF: {
                const T s = Fm<T>::sqrt(f->shelf);
                const T tsa = two * Fm<T>::sqrt(s) * alpha, sp1 = s + one, sm1 = s - one;
                o[0] = s * g * (sp1 + sm1 * fcos + tsa);
                o[1] = m2 * s * g * (sm1 + sp1 * fcos);
                o[2] = s * g * (sp1 + sm1 * fcos - tsa);
                o[3] = sp1 - sm1 * fcos + tsa;
                o[4] = two * (sm1 - sp1 * fcos);
                o[5] = sp1 - sm1 * fcos - tsa;
                return IDSP_OK;
            }
            case IDSP_IHO: {
                const T fs = (T)0.5 * Fm<T>::sin(f->frequency);
                const T a = (one + fcos) / (two * f->shelf);
                o[0] = g * (one + alpha); o[1] = m2 * g * fcos; o[2] = g * (one - alpha);
                o[3] = a + fs; o[4] = m2 * a; o[5] = a - fs;
                return IDSP_OK;
            }
        }
        return idsp_set_error("filter type must be 0..8 (coefficients::Type)"), IDSP_EINVAL;
    }
};

// biquad.rs:545-566
template <class T> void normalize(const T *ba6, T *n5) {
    const T a0 = (T)1.0 / ba6[3];
    n5[0] = ba6[0] * a0;
    n5[1] = ba6[1] * a0;
    n5[2] = ba6[2] * a0;
    n5[3] = -ba6[4] * a0;
    n5[4] = -ba6[5] * a0;
}

template <class T> int from_ba5(const T *n5, int kind, int F, void *out) {
    if (!n5 || !out) return idsp_set_error("null pointer argument"), IDSP_EINVAL;
    if (!kind_ok(kind, F)) return idsp_set_error("bad coefficient kind / F"), IDSP_EINVAL;
    for (int i = 0; i < 5; i++) put<T>(out, i, kind, F, n5[i]);
    return IDSP_OK;
}
template <class T> int from_ba6(const T *ba6, int kind, int F, void *out) {
    if (!ba6) return idsp_set_error("null pointer argument"), IDSP_EINVAL;
    T n5[5];
    normalize<T>(ba6, n5);
    return from_ba5<T>(n5, kind, F, out);
}

// pid.rs:193-222
template <class T, class PT> int pid_validate(const PT *b, T period) {
    if (!b) return idsp_set_error("null pointer argument"), IDSP_EINVAL;
    if (b->order < 0 || b->order > 2) return idsp_set_error("order must be 0 (I2), 1 (I) or 2 (P)"), IDSP_EINVAL;
    if (!std::isfinite(period)) return idsp_set_error("NonFinite(period)"), IDSP_EINVAL;
    if (period <= (T)0) return idsp_set_error("NonPositive(period)"), IDSP_EINVAL;
    for (int i = 0; i < 5; i++)
        if (b->gain[i] != b->gain[i]) return idsp_set_error("NonFinite(gain)"), IDSP_EINVAL;
    for (int i = 0; i < 5; i++)
        if (b->limit[i] != b->limit[i]) return idsp_set_error("NonFinite(limit)"), IDSP_EINVAL;
    const int actions[4] = {0, 1, 3, 4};  // I2, I, D, D2
    for (int a : actions) {
        const T gain = b->gain[a], limit = b->limit[a];
        if (std::isfinite(limit)) {
            if (limit == (T)0) return idsp_set_error("NonPositive(limit)"), IDSP_EINVAL;
            // Float::signum: 1.0 for +0.0 and positives, -1.0 for -0.0 and negatives
            if (gain != (T)0 && std::signbit(gain) != std::signbit(limit))
                return idsp_set_error("SignMismatch(gain/limit)"), IDSP_EINVAL;
        }
    }
    return IDSP_OK;
}

// coefficient arithmetic in the coefficient type C: `+=` / `-=` (wrapping for the Q formats)
template <class C> struct Acc {
    static C add(C a, C b) {
        if constexpr (std::is_integral<C>::value) {
            using U = typename std::make_unsigned<C>::type;
            return (C)(U)((U)a + (U)b);
        } else {
            return a + b;
        }
    }
    static C sub(C a, C b) {
        if constexpr (std::is_integral<C>::value) {
            using U = typename std::make_unsigned<C>::type;
            return (C)(U)((U)a - (U)b);
        } else {
            return a - b;
        }
    }
};
template <class C, class T> C to_c(T v, int F) {
    if constexpr (std::is_integral<C>::value) return quantize<C, T>(v, F);
    else return (C)v;
}

// pid.rs:236-303
template <class T, class PT, class C> void pid_build_c(const PT *b, T period, int F, C *out) {
    const int order = b->order;
    // period.powi(-order): llvm.powi = repeated multiplication, reciprocal last
    T z = order == 0 ? (T)1 : order == 1 ? (T)1 / period : (T)1 / (period * period);
    T gl[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    const int n = (5 - order) < 3 ? (5 - order) : 3;
    for (int j = n - 1; j >= 0; j--) {  // zip(...).rev(): highest action first
        const int i = order + j;
        gl[j][0] = b->gain[i] * z;
        gl[j][1] = i == 2 ? (T)1 : gl[j][0] / b->limit[i];
        z = z * period;
    }
    const T a0i = (T)1 / (gl[0][1] + gl[1][1] + gl[2][1]);
    const int kernels[3][3] = {{1, 0, 0}, {1, -1, 0}, {1, -2, 1}};
    C ba[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int g = 0; g < 3; g++) {
        const C g0 = to_c<C, T>(gl[g][0] * a0i, F), g1 = to_c<C, T>(gl[g][1] * a0i, F);  // quantise the gains
        for (int j = 0; j < 3; j++) {
            const int k = kernels[g][j];
            for (int r = 0; r < (k > 0 ? k : -k); r++) {
                if (k > 0) {
                    ba[j][0] = Acc<C>::add(ba[j][0], g0);
                    ba[j][1] = Acc<C>::sub(ba[j][1], g1);
                } else {
                    ba[j][0] = Acc<C>::sub(ba[j][0], g0);
                    ba[j][1] = Acc<C>::add(ba[j][1], g1);
                }
            }
        }
    }
    out[0] = ba[0][0];
    out[1] = ba[1][0];
    out[2] = ba[2][0];
    out[3] = ba[1][1];
    out[4] = ba[2][1];
}
template <class T, class PT> int pid_build(const PT *b, T period, int kind, int F, void *out) {
    if (!b || !out) return idsp_set_error("null pointer argument"), IDSP_EINVAL;
    if (b->order < 0 || b->order > 2) return idsp_set_error("order must be 0 (I2), 1 (I) or 2 (P)"), IDSP_EINVAL;
    if (!kind_ok(kind, F)) return idsp_set_error("bad coefficient kind / F"), IDSP_EINVAL;
    switch (kind) {
        case IDSP_I8: pid_build_c<T, PT, int8_t>(b, period, F, (int8_t *)out); break;
        case IDSP_I16: pid_build_c<T, PT, int16_t>(b, period, F, (int16_t *)out); break;
        case IDSP_I32: pid_build_c<T, PT, int32_t>(b, period, F, (int32_t *)out); break;
        case IDSP_I64: pid_build_c<T, PT, int64_t>(b, period, F, (int64_t *)out); break;
        case IDSP_F32: pid_build_c<T, PT, float>(b, period, F, (float *)out); break;
        default: pid_build_c<T, PT, double>(b, period, F, (double *)out); break;
    }
    return IDSP_OK;
}

}  // namespace

#define DEF_FLAVOUR(S, T)                                                                                     \
    extern "C" void idsp_filter_default_##S(idsp_filter_##S *f) {                                             \
        if (!f) return;                                                                                       \
        f->frequency = (T)0;                                                                                  \
        f->gain = (T)1;                                                                                       \
        f->shelf = (T)1;                                                                                      \
        f->shape_kind = IDSP_SHAPE_Q;                                                                         \
        f->shape = (T)1 / Fm<T>::SQRT_2; /* Shape::default() = Q(SQRT_2.recip()), coefficients.rs:18-22 */    \
    }                                                                                                         \
    extern "C" int idsp_filter_validate_##S(const idsp_filter_##S *f) {                                       \
        if (!f) return idsp_set_error("null pointer argument"), IDSP_EINVAL;                                  \
        return FilterImpl<T, idsp_filter_##S>::validate(f);                                                   \
    }                                                                                                         \
    extern "C" int idsp_filter_build_##S(const idsp_filter_##S *f, int type, T ba6[6]) {                      \
        if (!f || !ba6) return idsp_set_error("null pointer argument"), IDSP_EINVAL;                          \
        return FilterImpl<T, idsp_filter_##S>::build(f, type, ba6);                                           \
    }                                                                                                         \
    extern "C" int idsp_biquad_from_ba6_##S(const T ba6[6], int kind, int F, void *out) {                     \
        return from_ba6<T>(ba6, kind, F, out);                                                                \
    }                                                                                                         \
    extern "C" int idsp_biquad_from_ba5_##S(const T ba5[5], int kind, int F, void *out) {                     \
        return from_ba5<T>(ba5, kind, F, out);                                                                \
    }                                                                                                         \
    extern "C" int idsp_filter_build_biquad_##S(const idsp_filter_##S *f, int type, int kind, int F, void *out) { \
        int r = idsp_filter_validate_##S(f);                                                                  \
        if (r) return r;                                                                                      \
        T ba6[6];                                                                                             \
        r = idsp_filter_build_##S(f, type, ba6);                                                              \
        if (r) return r;                                                                                      \
        return from_ba6<T>(ba6, kind, F, out);                                                                \
    }                                                                                                         \
    extern "C" void idsp_pid_default_##S(idsp_pid_##S *b) {                                                   \
        if (!b) return;                                                                                       \
        b->order = 1;                                                                                         \
        for (int i = 0; i < 5; i++) {                                                                         \
            b->gain[i] = (T)0;                                                                                \
            b->limit[i] = std::numeric_limits<T>::infinity();                                                 \
        }                                                                                                     \
    }                                                                                                         \
    extern "C" int idsp_pid_validate_##S(const idsp_pid_##S *b, T period) {                                   \
        return pid_validate<T, idsp_pid_##S>(b, period);                                                      \
    }                                                                                                         \
    extern "C" int idsp_pid_build_##S(const idsp_pid_##S *b, T period, int kind, int F, void *out) {          \
        return pid_build<T, idsp_pid_##S>(b, period, kind, F, out);                                           \
    }
DEF_FLAVOUR(f64, double)
DEF_FLAVOUR(f32, float)

// biquad.rs:578-619
extern "C" int idsp_biquad_from_zpk_f64(const double zeros[2], int zeros_complex, const double poles[2],
                                        int poles_complex, double gain, int kind, int F, void *out) {
    if (!zeros || !poles) return idsp_set_error("null pointer argument"), IDSP_EINVAL;
    auto coeff = [](const double *p, int cplx, double *c) {
        if (cplx) {
            c[0] = p[0] + p[0];
            c[1] = p[0] * p[0] + p[1] * p[1];
        } else {
            c[0] = p[0] + p[1];
            c[1] = p[0] * p[1];
        }
    };
    double b[2], a[2];
    coeff(zeros, zeros_complex, b);
    coeff(poles, poles_complex, a);
    b[0] = gain * b[0];
    b[1] = gain * b[1];
    const double n5[5] = {gain, -b[0], b[1], a[0], -a[1]};
    return from_ba5<double>(n5, kind, F, out);
}
