// Compiled-host-side parity test: the reference's own known answers through the C++ mirror
// (include/idsp_b200.hpp) -> C ABI -> CUDA kernels.  Built and run by tests/test_gpu_cpp.py.
#include <cstdio>
#include <vector>

#include "idsp_b200.hpp"

using namespace idsp_b200;

#define EXPECT(c)                                                   \
    do {                                                            \
        if (!(c)) {                                                 \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                               \
        }                                                           \
    } while (0)

int main() {
    Engine e(0);
    {   // src/iir/coefficients.rs:289-301 (raw Q30 coefficients of the doctest's lowpass)
        Biquad<Q32<30>> iir{{Q32<30>::from_bits(2147483647), Q32<30>::from_bits(2147483647), Q32<30>::from_bits(2147483647),
                             Q32<30>::from_bits(1227265970), Q32<30>::from_bits(-443242341)}};
        std::vector<int32_t> xy = {3, -4, 5, 7, -3, 2};
        DirectForm1Lanes<int32_t> st(1);
        inplace(e, Lanes<Biquad<Q32<30>>>{iir}, st, xy.data(), xy.size());
        const int32_t want[6] = {5, 3, 9, 25, 42, 49};
        for (int i = 0; i < 6; i++) EXPECT(xy[i] == want[i]);
        EXPECT(st.x(0, 0) == 2 && st.x(0, 1) == -3 && st.y(0, 0) == 49 && st.y(0, 1) == 42);
    }
    {   // float -> Q (num_traits_impl.rs:32-45): saturation and half-away-from-zero
        EXPECT(Q32<30>::from_f64(2.0).bits == 2147483647);
        EXPECT(Q32<30>::from_f64(-2.0).bits == (int32_t)0x80000000);
        EXPECT((Q<int32_t, 0>::from_f64(0.5).bits == 1) && (Q<int32_t, 0>::from_f64(-0.5).bits == -1));
    }
    {   // src/iir/biquad.rs:130-155: offset and limits of BiquadClamp (zero coefficients)
        BiquadClamp<Q32<30>, int32_t> c{};
        for (auto &b : c.coeff.ba) b = Q32<30>::from_bits(0);
        c.u = 5;
        std::vector<int32_t> x = {0}, y = {0};
        DirectForm1Lanes<int32_t> st(1);
        block(e, Lanes<BiquadClamp<Q32<30>, int32_t>>{c}, st, x.data(), y.data(), 1);
        EXPECT(y[0] == 5);
        c.u = 0; c.max = -5;
        DirectForm1Lanes<int32_t> st2(1);
        block(e, Lanes<BiquadClamp<Q32<30>, int32_t>>{c}, st2, x.data(), y.data(), 1);
        EXPECT(y[0] == -5);
    }
    {   // Lanes: frame-major block == lane-major process_view on the transposed data
        const size_t lanes = 70, frames = 96;
        Biquad<Q32<30>> iir{{Q32<30>::from_f64(0.02), Q32<30>::from_f64(0.04), Q32<30>::from_f64(0.02),
                             Q32<30>::from_f64(1.5), Q32<30>::from_f64(-0.58)}};
        std::vector<int32_t> x(lanes * frames), xt(lanes * frames), y(lanes * frames), yt(lanes * frames);
        uint32_t s = 1;
        for (size_t t = 0; t < frames; t++)
            for (size_t l = 0; l < lanes; l++) {
                s = s * 1664525u + 1013904223u;
                x[t * lanes + l] = (int32_t)(s >> 4) - (1 << 27);
                xt[l * frames + t] = x[t * lanes + l];
            }
        DirectForm1Lanes<int32_t> a(lanes), b(lanes);
        block(e, Lanes<Biquad<Q32<30>>>{iir}, a, x.data(), y.data(), x.size());
        process_view(e, Lanes<Biquad<Q32<30>>>{iir}, b, View<const int32_t, LaneMajor>::from_flat(xt.data(), xt.size(), frames, lanes),
                     View<int32_t, LaneMajor>::from_flat(yt.data(), yt.size(), frames, lanes));
        for (size_t t = 0; t < frames; t++)
            for (size_t l = 0; l < lanes; l++) EXPECT(y[t * lanes + l] == yt[l * frames + t]);
        EXPECT(a.words == b.words);
        bool threw = false;
        try { View<int32_t, LaneMajor>::from_flat(yt.data(), yt.size() - 1, frames, lanes); } catch (const Error &) { threw = true; }
        EXPECT(threw);  // view.rs:181-182
    }
    {   // src/hbf.rs:576-595: /16 cascade response length = hbf_dec_response_length(4) = 57
        HbfDecState<4> h(1);
        std::vector<float> x(100 << 4), y(100);
        uint32_t s = 7;
        for (auto &v : x) { s = s * 1664525u + 1013904223u; v = (float)(s >> 8) / (float)(1 << 24); }
        block(e, Lanes<HbfDecCascade<4>>{}, h, x.data(), x.size(), y.data(), y.size());
        std::vector<float> z(1 << 10, 0.f), yz(1 << 6);
        block(e, Lanes<HbfDecCascade<4>>{}, h, z.data(), z.size(), yz.data(), yz.size());
        EXPECT(yz[56] != 0.0f);
        EXPECT(yz[57] == 0.0f);
    }
    {   // src/cossin.rs / src/atan2.rs exact values
        int32_t ph[2] = {0, 1 << 30}, cs[4];
        cossin(e, ph, cs, 2);
        EXPECT(cs[0] == 2147454703 && cs[1] == -1898);
        int32_t xy[8] = {1, 0, 2147483647, 0, 0, 1, 0, 2147483647}, p[4];
        atan2(e, xy, p, 4);
        EXPECT(p[0] == 0 && p[1] == 0 && p[2] == 0x3fffffff && p[3] == 0x3fffffff);
    }
    std::printf("cpp mirror ok\n");
    return 0;
}
