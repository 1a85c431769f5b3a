// hbf_stages.cuh -- half-band stage arithmetic shared by the HBF kernels.
//
// Arithmetic contract (src/hbf.rs:46-68, 155-236): for a window w[0..2M)
//   acc = ((w[2M-1]+w[0])*c[0]) + ((w[2M-2]+w[1])*c[1]) + ... (sequential, small taps first)
// every add/mul individually rounded (library built with -fmad=false), decimator
// output = acc + even sample, interpolator odd output = centre sample.
#pragma once
#include <stdint.h>

namespace idsp {

// HBF_TAPS (src/hbf.rs:308-349): filter design data of the reference, index 0 = lowest rate.
#define IDSP_HBF_TAPS0 {7.60375795e-07f, -3.77494111e-06f, 1.26458559e-05f, -3.43188253e-05f, \
    8.10687478e-05f, -1.72971467e-04f, 3.40845059e-04f, -6.29522864e-04f, 1.10128831e-03f,   \
    -1.83933299e-03f, 2.95124926e-03f, -4.57290964e-03f, 6.87374176e-03f, -1.00656257e-02f,  \
    1.44199840e-02f, -2.03025100e-02f, 2.82462332e-02f, -3.91128509e-02f, 5.44795658e-02f,   \
    -7.77002672e-02f, 1.17523452e-01f, -2.06185388e-01f, 6.34588695e-01f}
#define IDSP_HBF_TAPS1 {-1.12811343e-05f, 1.12724671e-04f, -6.07439343e-04f, 2.31904511e-03f, \
    -7.00322950e-03f, 1.78225473e-02f, -4.01209836e-02f, 8.43315989e-02f, -1.83189521e-01f,  \
    6.26346521e-01f}
#define IDSP_HBF_TAPS2 {0.0007686f, -0.00768669f, 0.0386536f, -0.14002434f, 0.60828885f}
#define IDSP_HBF_TAPS3 {-0.00261331f, 0.02476858f, -0.12112638f, 0.59897111f}
#define IDSP_HBF_TAPS4 {0.01186105f, -0.09808109f, 0.58622005f}

template <int IDX> struct HbfTaps;
template <> struct HbfTaps<0> { static constexpr int M = 23; __device__ __forceinline__ static float c(int i) { constexpr float t[23] = IDSP_HBF_TAPS0; return t[i]; } };
template <> struct HbfTaps<1> { static constexpr int M = 10; __device__ __forceinline__ static float c(int i) { constexpr float t[10] = IDSP_HBF_TAPS1; return t[i]; } };
template <> struct HbfTaps<2> { static constexpr int M = 5; __device__ __forceinline__ static float c(int i) { constexpr float t[5] = IDSP_HBF_TAPS2; return t[i]; } };
template <> struct HbfTaps<3> { static constexpr int M = 4; __device__ __forceinline__ static float c(int i) { constexpr float t[4] = IDSP_HBF_TAPS3; return t[i]; } };
template <> struct HbfTaps<4> { static constexpr int M = 3; __device__ __forceinline__ static float c(int i) { constexpr float t[3] = IDSP_HBF_TAPS4; return t[i]; } };

// HBF_TAPS_98 (src/hbf.rs:258-292): filter design data of the reference, index 0 = lowest rate.
#define IDSP_HBF98_TAPS0 {7.02144012e-05f, -2.43279582e-04f, 6.35026936e-04f, -1.39782541e-03f, 2.74613582e-03f, \
    -4.96403839e-03f, 8.41806912e-03f, -1.35827601e-02f, 2.11004053e-02f, -3.19267647e-02f, 4.77024289e-02f,   \
    -7.18014345e-02f, 1.12942004e-01f, -2.03279594e-01f, 6.33592923e-01f}
#define IDSP_HBF98_TAPS1 {-0.00086943f, 0.00577837f, -0.02201674f, 0.06357869f, -0.16627679f, 0.61979312f}
#define IDSP_HBF98_TAPS2 {0.01414651f, -0.10439639f, 0.59026742f}
#define IDSP_HBF98_TAPS3 {0.01227974f, -0.09930782f, 0.58702834f}
#define IDSP_HBF98_TAPS4 {-0.06291796f, 0.5629161f}
template <int IDX> struct HbfTaps98;
template <> struct HbfTaps98<0> { static constexpr int M = 15; __device__ __forceinline__ static float c(int i) { constexpr float t[15] = IDSP_HBF98_TAPS0; return t[i]; } };
template <> struct HbfTaps98<1> { static constexpr int M = 6; __device__ __forceinline__ static float c(int i) { constexpr float t[6] = IDSP_HBF98_TAPS1; return t[i]; } };
template <> struct HbfTaps98<2> { static constexpr int M = 3; __device__ __forceinline__ static float c(int i) { constexpr float t[3] = IDSP_HBF98_TAPS2; return t[i]; } };
template <> struct HbfTaps98<3> { static constexpr int M = 3; __device__ __forceinline__ static float c(int i) { constexpr float t[3] = IDSP_HBF98_TAPS3; return t[i]; } };
template <> struct HbfTaps98<4> { static constexpr int M = 2; __device__ __forceinline__ static float c(int i) { constexpr float t[2] = IDSP_HBF98_TAPS4; return t[i]; } };
__host__ __device__ constexpr int hbf98_m(int idx) {
    return idx == 0 ? 15 : idx == 1 ? 6 : idx == 2 ? 3 : idx == 3 ? 3 : 2;
}

__host__ __device__ constexpr int hbf_m(int idx) {
    return idx == 0 ? 23 : idx == 1 ? 10 : idx == 2 ? 5 : idx == 3 ? 4 : 3;
}
__host__ __device__ constexpr int hbf_dec_words(int k) {  // sum of (M-1)+(2M-1)
    int w = 0;
    for (int i = 0; i < k; i++) w += 3 * hbf_m(i) - 2;
    return w;
}
__host__ __device__ constexpr int hbf_int_words(int k) {
    int w = 0;
    for (int i = 0; i < k; i++) w += 2 * hbf_m(i) - 1;
    return w;
}

// One /2 stage with its delay lines in registers (shift-register form; used by the
// generic thread-per-lane kernels).  State word order = ABI: even(M-1) | odd(2M-1).
template <int IDX> struct DecStageRegs {
    static constexpr int M = HbfTaps<IDX>::M;
    static constexpr int LEN = 2 * M - 1;
    static constexpr int WORDS = 3 * M - 2;
    float ev[M - 1];
    float od[LEN];
    __device__ __forceinline__ void load(const float *st, size_t stride, size_t lane) {
#pragma unroll
        for (int i = 0; i < M - 1; i++) ev[i] = st[(size_t)i * stride + lane];
#pragma unroll
        for (int i = 0; i < LEN; i++) od[i] = st[(size_t)(M - 1 + i) * stride + lane];
    }
    __device__ __forceinline__ void store(float *st, size_t stride, size_t lane) const {
#pragma unroll
        for (int i = 0; i < M - 1; i++) st[(size_t)i * stride + lane] = ev[i];
#pragma unroll
        for (int i = 0; i < LEN; i++) st[(size_t)(M - 1 + i) * stride + lane] = od[i];
    }
    __device__ __forceinline__ float push(float e, float o) {
        float acc = (o + od[0]) * HbfTaps<IDX>::c(0);
#pragma unroll
        for (int i = 1; i < M; i++) acc = acc + (od[LEN - i] + od[i]) * HbfTaps<IDX>::c(i);
        float y = acc + ev[0];
#pragma unroll
        for (int i = 0; i < LEN - 1; i++) od[i] = od[i + 1];
        od[LEN - 1] = o;
#pragma unroll
        for (int i = 0; i < M - 2; i++) ev[i] = ev[i + 1];
        ev[M - 2] = e;
        return y;
    }
};

// One x2 stage: state x history (2M-1).
template <int IDX> struct IntStageRegs {
    static constexpr int M = HbfTaps<IDX>::M;
    static constexpr int LEN = 2 * M - 1;
    static constexpr int WORDS = LEN;
    float xs[LEN];
    __device__ __forceinline__ void load(const float *st, size_t stride, size_t lane) {
#pragma unroll
        for (int i = 0; i < LEN; i++) xs[i] = st[(size_t)i * stride + lane];
    }
    __device__ __forceinline__ void store(float *st, size_t stride, size_t lane) const {
#pragma unroll
        for (int i = 0; i < LEN; i++) st[(size_t)i * stride + lane] = xs[i];
    }
    // returns (even = interpolated, odd = centre tap)
    __device__ __forceinline__ float2 push(float x) {
        float acc = (x + xs[0]) * HbfTaps<IDX>::c(0);
#pragma unroll
        for (int i = 1; i < M; i++) acc = acc + (xs[LEN - i] + xs[i]) * HbfTaps<IDX>::c(i);
        // window w = [xs..., x]; centre = w[M] (hbf.rs:222: state.x[M..])
        float odd = (M == LEN) ? x : xs[M];
#pragma unroll
        for (int i = 0; i < LEN - 1; i++) xs[i] = xs[i + 1];
        xs[LEN - 1] = x;
        return make_float2(acc, odd);
    }
};

}  // namespace idsp
