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
    T u = 0, min = std::numeric_limits<T>::lowest(), max = std::numeric_limits<T>::max();
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

// cossin(phase) -> (cos, sin), atan2(y, x) (src/cossin.rs:14-67, src/atan2.rs:66-82), slice forms
inline void cossin(Engine &e, const int32_t *phase, int32_t *cs, size_t n) {
    check(idsp_cossin_i32_host(e.ctx(), phase, cs, n));
}
inline void atan2(Engine &e, const int32_t *xy, int32_t *p, size_t n) {
    check(idsp_atan2_i32_host(e.ctx(), xy, p, n));
}

}  // namespace idsp_b200
