// Compiled-host-side parity test: the reference's own known answers through the C++ mirror
// (include/idsp_b200.hpp) -> C ABI -> CUDA kernels.  Built and run by tests/test_gpu_cpp.py.
#include <cstdio>
#include <cstring>
#include <string>
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

static bool dump(const std::string &path, const void *p, size_t bytes) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(p, 1, bytes, f) == bytes;
    std::fclose(f);
    return ok;
}

int main(int argc, char **argv) {
    Engine e(0);
    const std::string outdir = argc > 1 ? argv[1] : "";
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
    {   // device-resident graph (compose.rs:13-113): HbfDec16 -> HbfInt16 -> Biquad<f32> on GpuBuffers.
        // One upload, one download; fused single entry point == the three stages through device
        // intermediates, streamed in two calls (state carried); tests/test_gpu_cpp.py compares the dump
        // with the CPU oracle bit for bit.
        constexpr int K = 4;
        const size_t lanes = 48, n_low = 96, len = lanes * (n_low << K);
        std::vector<float> x(len), y1(len), y2(len);
        uint32_t s = 99;
        for (auto &v : x) { s = s * 1664525u + 1013904223u; v = (float)(int32_t)s / 2147483648.0f; }
        Biquad<float> iir = Filter().critical_frequency(0.05).build_biquad_f32(IDSP_LOWPASS);
        Lanes<DecIntBiquad<K>> graph{{iir}};
        GpuBuffer<float> dx(e, len), dy(e, len);
        dx.upload(x.data(), len);
        DecIntBiquadState<K> st1(e, lanes), st2(e, lanes);
        block(e, graph, st1, dx, dy, len, true, LaneMajor{});
        dy.download(y1.data(), len);
        block(e, graph, st2, dx, dy, len, false, LaneMajor{});
        dy.download(y2.data(), len);
        for (size_t i = 0; i < len; i++) EXPECT(std::memcmp(&y1[i], &y2[i], 4) == 0);
        std::vector<float> w1(st1.words.size()), w2(st2.words.size());
        st1.words.download(w1.data(), w1.size());
        st2.words.download(w2.data(), w2.size());
        EXPECT(std::memcmp(w1.data(), w2.data(), w1.size() * 4) == 0);
        // the same graph on host slices through one PCIe round trip
        std::vector<float> hw(idsp_chain_state_words(K) * lanes, 0.f), y3(len);
        block(e, graph, hw, lanes, x.data(), y3.data(), len, LaneMajor{});
        EXPECT(std::memcmp(y1.data(), y3.data(), len * 4) == 0);
        EXPECT(std::memcmp(w1.data(), hw.data(), hw.size() * 4) == 0);
        if (!outdir.empty()) {
            EXPECT(dump(outdir + "/chain_x.bin", x.data(), len * 4));
            EXPECT(dump(outdir + "/chain_y.bin", y1.data(), len * 4));
            EXPECT(dump(outdir + "/chain_ba.bin", iir.ba.data(), 5 * 4));
            EXPECT(dump(outdir + "/chain_state.bin", w1.data(), w1.size() * 4));
        }
    }
    {   // device-resident lanes of the i32 biquad + lock-in (no PCIe inside the graph)
        const size_t lanes = 256, frames = 64, n = lanes * frames;
        Biquad<Q32<30>> iir = Filter().critical_frequency(0.1).gain(1000.0).build_biquad<30>(IDSP_LOWPASS);
        EXPECT(iir.ba[3].bits == 1227265970 && iir.ba[4].bits == -443242341 && iir.ba[0].bits == 2147483647);
        std::vector<int32_t> x(n), y(n), yh(n);
        uint32_t s = 5;
        for (auto &v : x) { s = s * 1664525u + 1013904223u; v = (int32_t)(s >> 4) - (1 << 27); }
        GpuBuffer<int32_t> dx(e, n), dy(e, n), diq(e, 2 * n);
        dx.upload(x.data(), n);
        auto st = df1_state_i32(e, lanes);
        block(e, Lanes<Biquad<Q32<30>>>{iir}, st, dx, dy, n);
        dy.download(y.data(), n);
        DirectForm1Lanes<int32_t> hs(lanes);
        block(e, Lanes<Biquad<Q32<30>>>{iir}, hs, x.data(), yh.data(), n);  // host-slice path
        EXPECT(y == yh);
        LockinState<2> ls(e, lanes);
        std::vector<int32_t> step(lanes);
        for (size_t l = 0; l < lanes; l++) step[l] = (int32_t)(0x01234567u * (uint32_t)(l + 1));
        ls.accu_step.upload(step.data(), lanes);
        Lanes<Lockin<Lowpass<2>>> lk{{{{1048576, -94906265}}}};
        block(e, lk, ls, dy, diq, n);  // the biquad's output feeds the lock-in without leaving the device
        std::vector<int32_t> iq(2 * n), acc(lanes);
        diq.download(iq.data(), 2 * n);
        ls.accu_state.download(acc.data(), lanes);
        for (size_t l = 0; l < lanes; l++) EXPECT(acc[l] == (int32_t)((uint32_t)step[l] * (uint32_t)frames));  // accu.rs:34-37
        if (!outdir.empty()) {
            EXPECT(dump(outdir + "/lockin_x.bin", y.data(), n * 4));
            EXPECT(dump(outdir + "/lockin_step.bin", step.data(), lanes * 4));
            EXPECT(dump(outdir + "/lockin_iq.bin", iq.data(), 2 * n * 4));
        }
    }
    {   // validation errors of the builders cross the ABI as IDSP_EINVAL + message
        bool threw = false;
        try { Filter().critical_frequency(0.7).validate(); } catch (const Error &err) { threw = std::string(err.what()).find("OutOfRange(frequency)") != std::string::npos; }
        EXPECT(threw);
    }
    std::printf("cpp mirror ok\n");
    return 0;
}
