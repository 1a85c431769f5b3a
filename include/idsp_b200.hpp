// idsp_b200.hpp -- header-only C++17 host mirror of the reference's operator surface on
// top of the C ABI (idsp_b200.h).  The reference is compiled Rust and its toolchain is not
// available in the build image, so this is the compiled-language host side: same names,
// argument meaning and error behaviour as
//   dsp_process::{SplitProcess::block, Split, Lanes, View}   (dsp-process/src/*.rs)
//   idsp::iir::{Biquad, BiquadClamp, DirectForm1}            (src/iir/biquad.rs)
//   idsp::hbf::{HBF_DEC_CASCADE, HbfDec2..32}                (src/hbf.rs)
//   idsp::{cossin, atan2}                                    (src/cossin.rs, src/atan2.rs)
// Host slices go through the `*_host` entry points (the library streams them through the
// GPU); lane count is a run-time property of the state (the reference's `[S; N]` const
// generic cannot hold 2^16..2^24 lanes).
#pragma once
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "idsp_b200.h"

namespace idsp_b200 {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc) {
    if (rc != 0) throw Error(std::string("idsp_b200: ") + idsp_b200_last_error());
}

// One device + one stream; not thread-safe (mirrors `&mut` exclusivity of the states).
class Engine {
   public:
    explicit Engine(int device = 0) { check(idsp_b200_init(device, &ctx_)); }
    ~Engine() { idsp_b200_free(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    idsp_ctx *ctx() const { return ctx_; }
    void sync() { check(idsp_b200_sync(ctx_)); }

   private:
    idsp_ctx *ctx_ = nullptr;
};

struct FrameMajor { static constexpr int value = IDSP_FRAME_MAJOR; };
struct LaneMajor { static constexpr int value = IDSP_LANE_MAJOR; };

// dsp-process/src/view.rs:24-36: typed view of a flat slice; from_flat asserts the length
// (view.rs:181-182 panics there, throws here).
template <class T, class Layout> struct View {
    T *flat;
    size_t frames, lanes;
    static View from_flat(T *flat, size_t len, size_t frames, size_t lanes) {
        if (len != frames * lanes) throw Error("View::from_flat: flat.len() != frames * L");
        return View{flat, frames, lanes};
    }
};

// Q<T, A, F> (dsp-fixedpoint/src/lib.rs:155-160): raw bits + float conversion
// `(v * 2^F).round() as T` = half away from zero, saturating, NaN -> 0.
template <class T, int F> struct Q {
    T bits;
    static Q from_bits(T b) { return Q{b}; }
    static Q from_f64(double v) {
        double s = std::ldexp(v, F);
        if (s != s) return Q{0};
        double r = std::round(s);
        if (r >= (double)std::numeric_limits<T>::max()) return Q{std::numeric_limits<T>::max()};
        if (r <= (double)std::numeric_limits<T>::min()) return Q{std::numeric_limits<T>::min()};
        return Q{(T)r};
    }
};
template <int F> using Q32 = Q<int32_t, F>;

// Biquad<C>: ba = [b0, b1, b2, a1, a2] (src/iir/biquad.rs:96-116)
template <class C> struct Biquad {
    std::array<C, 5> ba;
};
// From<[[f64;3];2]> (src/iir/biquad.rs:545-566): literature signs -> normalised, sign flipped
template <int F> Biquad<Q32<F>> biquad_from_ba6(const double (&b)[3], const double (&a)[3]) {
    double a0 = 1.0 / a[0];
    return Biquad<Q32<F>>{{Q32<F>::from_f64(b[0] * a0), Q32<F>::from_f64(b[1] * a0), Q32<F>::from_f64(b[2] * a0),
                           Q32<F>::from_f64(-a[1] * a0), Q32<F>::from_f64(-a[2] * a0)}};
}
// BiquadClamp<C, T> (src/iir/biquad.rs:121-171)
template <class C, class T> struct BiquadClamp {
    Biquad<C> coeff;
    // Clamp::MIN / MAX (src/num.rs:33-54): -inf / +inf for floats, the integer limits otherwise
    T u = 0;
    T min = std::numeric_limits<T>::has_infinity ? -std::numeric_limits<T>::infinity() : std::numeric_limits<T>::lowest();
    T max = std::numeric_limits<T>::has_infinity ? std::numeric_limits<T>::infinity() : std::numeric_limits<T>::max();
};

// [DirectForm1<T>; N] as the ABI's SoA words [x0 | x1 | y0 | y1] x lanes (biquad.rs:260-269)
template <class T> struct DirectForm1Lanes {
    size_t lanes;
    std::vector<T> words;
    explicit DirectForm1Lanes(size_t n) : lanes(n), words(4 * n, T(0)) {}
    T x(size_t lane, int i) const { return words[i * lanes + lane]; }
    T y(size_t lane, int i) const { return words[(2 + i) * lanes + lane]; }
    void set_y(T v) {  // biquad.rs:296-300
        for (size_t l = 0; l < 2 * lanes; l++) words[2 * lanes + l] = v;
    }
};

// Lanes<C> (dsp-process/src/compose.rs:448-513)
template <class C> struct Lanes {
    C inner;
};

// SplitProcess::block for Lanes<Biquad<Q32<F>>> on [DirectForm1<i32>; N], frame-major
// `[[i32; N]]` (compose.rs:468-476) or lane-major views (compose.rs:478-494).
template <int F, class Layout = FrameMajor>
void block(Engine &e, const Lanes<Biquad<Q32<F>>> &c, DirectForm1Lanes<int32_t> &state, const int32_t *x,
           int32_t *y, size_t len, Layout = Layout{}) {
    if (len % state.lanes) throw Error("block: length is not a whole number of frames");
    int32_t ba[5];
    for (int i = 0; i < 5; i++) ba[i] = c.inner.ba[i].bits;
    check(idsp_biquad_df1_i32_host(e.ctx(), ba, F, nullptr, state.words.data(), x, y, len / state.lanes,
                                   state.lanes, Layout::value));
}
template <int F, class Layout = FrameMajor>
void block(Engine &e, const Lanes<BiquadClamp<Q32<F>, int32_t>> &c, DirectForm1Lanes<int32_t> &state,
           const int32_t *x, int32_t *y, size_t len, Layout = Layout{}) {
    if (len % state.lanes) throw Error("block: length is not a whole number of frames");
    int32_t ba[5], cl[3] = {c.inner.u, c.inner.min, c.inner.max};
    for (int i = 0; i < 5; i++) ba[i] = c.inner.coeff.ba[i].bits;
    check(idsp_biquad_df1_i32_host(e.ctx(), ba, F, cl, state.words.data(), x, y, len / state.lanes,
                                   state.lanes, Layout::value));
}
// SplitInplace::inplace (process.rs:135-142)
template <class C, class S> void inplace(Engine &e, const C &c, S &state, int32_t *xy, size_t len) {
    block(e, c, state, xy, xy, len);
}
// ViewProcess::process_view on lane-major views (view.rs:295-302)
template <int F>
void process_view(Engine &e, const Lanes<Biquad<Q32<F>>> &c, DirectForm1Lanes<int32_t> &state,
                  View<const int32_t, LaneMajor> x, View<int32_t, LaneMajor> y) {
    if (x.frames != y.frames || x.lanes != state.lanes || y.lanes != state.lanes)
        throw Error("process_view: shape mismatch");
    block(e, c, state, x.flat, y.flat, x.frames * x.lanes, LaneMajor{});
}

// HBF_DEC_CASCADE truncated to depth K with its HbfDec{2,4,8,16,32} state (hbf.rs:363-421)
template <int K> struct HbfDecState {
    size_t lanes;
    std::vector<float> words;
    explicit HbfDecState(size_t n) : lanes(n), words(idsp_hbf_dec_state_words(K) * n, 0.f) {}
};
template <int K> struct HbfDecCascade {};
template <int K, class Layout = FrameMajor>
void block(Engine &e, const Lanes<HbfDecCascade<K>> &, HbfDecState<K> &state, const float *x, size_t x_len,
           float *y, size_t y_len, Layout = Layout{}) {
    if (x_len != (y_len << K) || y_len % state.lanes) throw Error("block: x and y lengths do not match");
    check(idsp_hbf_dec_cascade_f32_host(e.ctx(), K, state.words.data(), x, y, y_len / state.lanes, state.lanes,
                                        Layout::value));
}

// ---------------------------------------------------------------------------------------------
// Device-resident buffers and states: a chained graph (the reference's tuple chaining,
// dsp-process/src/compose.rs:13-113, and `Major` scratch-buffered chaining :569-613) keeps its samples
// and filter states in HBM between stages and crosses PCIe once at each end -- this is what makes the
// resident-data throughput of the kernels reachable from a reference-side caller.
// ---------------------------------------------------------------------------------------------
template <class T> class GpuBuffer {
   public:
    GpuBuffer(Engine &e, size_t n, bool zero = false) : e_(&e), n_(n) {
        void *p = nullptr;
        check(idsp_b200_malloc(e.ctx(), n * sizeof(T), &p));
        p_ = static_cast<T *>(p);
        if (zero) check(idsp_b200_memset(e.ctx(), p_, 0, n * sizeof(T)));
    }
    ~GpuBuffer() {
        if (p_) idsp_b200_mfree(e_->ctx(), p_);
    }
    GpuBuffer(const GpuBuffer &) = delete;
    GpuBuffer &operator=(const GpuBuffer &) = delete;
    GpuBuffer(GpuBuffer &&o) noexcept : e_(o.e_), p_(o.p_), n_(o.n_) { o.p_ = nullptr; }
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t size() const { return n_; }
    void upload(const T *host, size_t n) {
        if (n > n_) throw Error("GpuBuffer::upload: too long");
        check(idsp_b200_memcpy(e_->ctx(), p_, host, n * sizeof(T), 0));
    }
    void download(T *host, size_t n) const {  // synchronises: the data is in `host` on return
        if (n > n_) throw Error("GpuBuffer::download: too long");
        check(idsp_b200_memcpy(e_->ctx(), host, p_, n * sizeof(T), 1));
        check(idsp_b200_sync(e_->ctx()));
    }

   private:
    Engine *e_;
    T *p_ = nullptr;
    size_t n_;
};
// `[S; N]` on the device: zeroed SoA words (= `Default`), word count from the ABI
template <class T> struct GpuState {
    size_t lanes;
    GpuBuffer<T> words;
    GpuState(Engine &e, size_t words_per_lane, size_t n) : lanes(n), words(e, words_per_lane * n, true) {}
};

// Lanes<Biquad<Q32<F>>> on [DirectForm1<i32>; N], device buffers (src/iir/biquad.rs:366-383)
template <int F, class Layout = FrameMajor>
void block(Engine &e, const Lanes<Biquad<Q32<F>>> &c, GpuState<int32_t> &state, const GpuBuffer<int32_t> &x,
           GpuBuffer<int32_t> &y, size_t len, Layout = Layout{}) {
    if (len % state.lanes || len > x.size() || len > y.size()) throw Error("block: bad length");
    int32_t ba[5];
    for (int i = 0; i < 5; i++) ba[i] = c.inner.ba[i].bits;
    check(idsp_biquad_df1_i32(e.ctx(), ba, F, nullptr, state.words.data(), x.data(), y.data(), len / state.lanes,
                              state.lanes, Layout::value));
}
// Lanes<Biquad<f32>> on [DirectForm1<f32>; N]
template <class Layout = FrameMajor>
void block(Engine &e, const Lanes<Biquad<float>> &c, GpuState<float> &state, const GpuBuffer<float> &x,
           GpuBuffer<float> &y, size_t len, Layout = Layout{}) {
    if (len % state.lanes || len > x.size() || len > y.size()) throw Error("block: bad length");
    check(idsp_biquad_df1_f32(e.ctx(), c.inner.ba.data(), 0, nullptr, state.words.data(), x.data(), y.data(),
                              len / state.lanes, state.lanes, Layout::value));
}
inline GpuState<int32_t> df1_state_i32(Engine &e, size_t lanes) { return GpuState<int32_t>(e, 4, lanes); }
inline GpuState<float> df1_state_f32(Engine &e, size_t lanes) { return GpuState<float>(e, 4, lanes); }

// HBF_DEC_CASCADE / HBF_INT_CASCADE truncated to depth K on device buffers (src/hbf.rs:385-421, 476-512):
// SplitProcess<[f32; 2^K], f32, HbfDec..> and SplitProcess<f32, [f32; 2^K], HbfInt..>
template <int K> struct HbfIntCascade {};
template <int K> GpuState<float> hbf_dec_state(Engine &e, size_t lanes) { return GpuState<float>(e, idsp_hbf_dec_state_words(K), lanes); }
template <int K> GpuState<float> hbf_int_state(Engine &e, size_t lanes) { return GpuState<float>(e, idsp_hbf_int_state_words(K), lanes); }
template <int K, class Layout = FrameMajor>
void block(Engine &e, const Lanes<HbfDecCascade<K>> &, GpuState<float> &state, const GpuBuffer<float> &x, size_t x_len,
           GpuBuffer<float> &y, size_t y_len, Layout = Layout{}) {
    if (x_len != (y_len << K) || y_len % state.lanes || x_len > x.size() || y_len > y.size())
        throw Error("block: x and y lengths do not match");
    check(idsp_hbf_dec_cascade_f32(e.ctx(), K, state.words.data(), x.data(), y.data(), y_len / state.lanes, state.lanes,
                                   Layout::value));
}
template <int K, class Layout = FrameMajor>
void block(Engine &e, const Lanes<HbfIntCascade<K>> &, GpuState<float> &state, const GpuBuffer<float> &x, size_t x_len,
           GpuBuffer<float> &y, size_t y_len, Layout = Layout{}) {
    if (y_len != (x_len << K) || x_len % state.lanes || x_len > x.size() || y_len > y.size())
        throw Error("block: x and y lengths do not match");
    check(idsp_hbf_int_cascade_f32(e.ctx(), K, state.words.data(), x.data(), y.data(), x_len / state.lanes, state.lanes,
                                   Layout::value));
}

// Tuple chaining that stays on the device: (HbfDecCascade<K>, HbfIntCascade<K>, Biquad<f32>) as one
// processor, like `(a, b, c)` / `a * b * c` in the reference (compose.rs:13-113).  `fused` runs the
// library's single entry point (idsp_chain_f32: one or two passes over HBM), otherwise the three stages
// run one after the other through device-resident intermediates; both give identical bits.
template <int K> struct DecIntBiquad {
    Biquad<float> iir;
};
template <int K> struct DecIntBiquadState {
    size_t lanes;
    GpuBuffer<float> words;  // [dec cascade | int cascade | DirectForm1] per lane, SoA
    DecIntBiquadState(Engine &e, size_t n) : lanes(n), words(e, idsp_chain_state_words(K) * n, true) {}
};
template <int K, class Layout = FrameMajor>
void block(Engine &e, const Lanes<DecIntBiquad<K>> &c, DecIntBiquadState<K> &state, const GpuBuffer<float> &x,
           GpuBuffer<float> &y, size_t len, bool fused = true, Layout = Layout{}) {
    const size_t lanes = state.lanes;
    if (len % (lanes << K) || len > x.size() || len > y.size()) throw Error("block: bad length");
    const size_t n_low = len / (lanes << K);
    if (fused) {
        check(idsp_chain_f32(e.ctx(), K, c.inner.iir.ba.data(), state.words.data(), x.data(), y.data(), n_low, lanes,
                             Layout::value));
        return;
    }
    GpuBuffer<float> low(e, n_low * lanes);
    float *st = state.words.data();
    const size_t wd = idsp_hbf_dec_state_words(K), wi = idsp_hbf_int_state_words(K);
    check(idsp_hbf_dec_cascade_f32(e.ctx(), K, st, x.data(), low.data(), n_low, lanes, Layout::value));
    check(idsp_hbf_int_cascade_f32(e.ctx(), K, st + wd * lanes, low.data(), y.data(), n_low, lanes, Layout::value));
    check(idsp_biquad_df1_f32(e.ctx(), c.inner.iir.ba.data(), 0, nullptr, st + (wd + wi) * lanes, y.data(), y.data(),
                              n_low << K, lanes, Layout::value));
    e.sync();  // `low` is released on return
}
// one PCIe round trip for the three operators on host slices
template <int K, class Layout = FrameMajor>
void block(Engine &e, const Lanes<DecIntBiquad<K>> &c, std::vector<float> &state_words, size_t lanes, const float *x,
           float *y, size_t len, Layout = Layout{}) {
    if (len % (lanes << K) || state_words.size() != idsp_chain_state_words(K) * lanes) throw Error("block: bad length");
    check(idsp_chain_f32_host(e.ctx(), K, c.inner.iir.ba.data(), state_words.data(), x, y, len / (lanes << K), lanes,
                              Layout::value));
}

// Lowpass<N> / Lockin<Lowpass<N>> with the phase from a per-lane Accu (src/lowpass.rs:13, src/lockin.rs:30-39,
// src/accu.rs:29-38), device buffers; iq = Complex<i32> per sample
template <int N> struct Lowpass {
    std::array<int32_t, N> k;
};
template <class C> struct Lockin {
    C lowpass;
};
template <int N> struct LockinState {
    size_t lanes;
    GpuBuffer<int32_t> accu_state, accu_step;
    GpuBuffer<int64_t> lp;
    LockinState(Engine &e, size_t n) : lanes(n), accu_state(e, n, true), accu_step(e, n, true), lp(e, 2 * N * n, true) {}
};
template <int N, class Layout = FrameMajor>
void block(Engine &e, const Lanes<Lockin<Lowpass<N>>> &c, LockinState<N> &state, const GpuBuffer<int32_t> &x,
           GpuBuffer<int32_t> &iq, size_t len, Layout = Layout{}) {
    if (len % state.lanes || len > x.size() || 2 * len > iq.size()) throw Error("block: bad length");
    check(idsp_lockin_i32(e.ctx(), N, c.inner.lowpass.k.data(), state.accu_state.data(), state.accu_step.data(),
                          state.lp.data(), x.data(), iq.data(), len / state.lanes, state.lanes, Layout::value));
}

// coefficient builders across the ABI (src/iir/coefficients.rs, src/iir/pid.rs): design parameters in,
// a Biquad<Q32<F>> / Biquad<f32> out
struct Filter {
    idsp_filter_f64 f;
    Filter() { idsp_filter_default_f64(&f); }
    Filter &critical_frequency(double f0) { f.frequency = 6.283185307179586476925286766559 * f0; return *this; }
    Filter &gain(double k) { f.gain = k; return *this; }
    Filter &shelf(double a) { f.shelf = a; return *this; }
    Filter &q(double v) { f.shape_kind = IDSP_SHAPE_Q; f.shape = v; return *this; }
    Filter &bandwidth(double v) { f.shape_kind = IDSP_SHAPE_BANDWIDTH; f.shape = v; return *this; }
    Filter &shelf_slope(double v) { f.shape_kind = IDSP_SHAPE_SLOPE; f.shape = v; return *this; }
    void validate() const { check(idsp_filter_validate_f64(&f)); }
    template <int F> Biquad<Q32<F>> build_biquad(idsp_filter_type_t typ) const {
        int32_t raw[5];
        check(idsp_filter_build_biquad_f64(&f, typ, IDSP_I32, F, raw));
        Biquad<Q32<F>> b;
        for (int i = 0; i < 5; i++) b.ba[i] = Q32<F>::from_bits(raw[i]);
        return b;
    }
    Biquad<float> build_biquad_f32(idsp_filter_type_t typ) const {
        Biquad<float> b;
        check(idsp_filter_build_biquad_f64(&f, typ, IDSP_F32, 0, b.ba.data()));
        return b;
    }
};

// cossin(phase) -> (cos, sin), atan2(y, x) (src/cossin.rs:14-67, src/atan2.rs:66-82), slice forms
inline void cossin(Engine &e, const int32_t *phase, int32_t *cs, size_t n) {
    check(idsp_cossin_i32_host(e.ctx(), phase, cs, n));
}
inline void atan2(Engine &e, const int32_t *xy, int32_t *p, size_t n) {
    check(idsp_atan2_i32_host(e.ctx(), xy, p, n));
}

}  // namespace idsp_b200
