// cic.cu -- Cic<T, N, M> decimator / interpolator lanes (include/idsp_b200.h).
//
// Reference: src/cic.rs:13-200 under the Decimator / Interpolator adapters of
// dsp-process/src/adapters.rs:27-35, :154-222.  One filter lane per thread: the N integrators
// run at the high rate and are a serial chain of wrapping adds per lane, the N combs (delay M)
// run once per frame.  State (SoA words of T): [index, zoh, combs[N][M], integrators[N]].
// A frame is R = rate + 1 consecutive high-rate samples of one lane, so a thread reads
// (decimator) or writes (interpolator) R contiguous elements per frame, 16 bytes at a time
// when R and the base pointer allow it.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int CIC_MAXM = 3;

template <class T> struct CicParams {
    T *st;
    uint32_t rate;
    int M;
};

template <class T, int N, int M> struct CicRegs {
    using UT = typename std::make_unsigned<T>::type;
    UT integ[N], comb[N][M], zoh;
    uint32_t index;
    __device__ __forceinline__ void load(const CicParams<T> &p, size_t lane, size_t stride) {
        index = (uint32_t)p.st[lane];
        zoh = (UT)p.st[stride + lane];
#pragma unroll
        for (int n = 0; n < N; n++) {
#pragma unroll
            for (int m = 0; m < M; m++) comb[n][m] = (UT)p.st[(size_t)(2 + n * M + m) * stride + lane];
            integ[n] = (UT)p.st[(size_t)(2 + N * M + n) * stride + lane];
        }
    }
    __device__ __forceinline__ void store(const CicParams<T> &p, size_t lane, size_t stride) const {
        p.st[lane] = (T)index;
        p.st[stride + lane] = (T)zoh;
#pragma unroll
        for (int n = 0; n < N; n++) {
#pragma unroll
            for (int m = 0; m < M; m++) p.st[(size_t)(2 + n * M + m) * stride + lane] = (T)comb[n][m];
            p.st[(size_t)(2 + N * M + n) * stride + lane] = (T)integ[n];
        }
    }
    // comb cascade (src/cic.rs:159-164, :190-196): y = x - c[0]; c shifts down; c[M-1] = x
    __device__ __forceinline__ UT combs(UT x) {
#pragma unroll
        for (int n = 0; n < N; n++) {
            const UT y = x - comb[n][0];
#pragma unroll
            for (int m = 0; m + 1 < M; m++) comb[n][m] = comb[n][m + 1];
            comb[n][M - 1] = x;
            x = y;
        }
        return x;
    }
    // Process<T, Option<T>> (src/cic.rs:176-200); the tick value is left in zoh
    __device__ __forceinline__ void dec_step(uint32_t rate, UT x) {
#pragma unroll
        for (int n = 0; n < N; n++) {
            integ[n] += x;
            x = integ[n];
        }
        if (index) {
            index--;
        } else {
            index = rate;
            zoh = combs(x);
        }
    }
    // Process<Option<T>, T> (src/cic.rs:149-172)
    __device__ __forceinline__ UT int_step(uint32_t rate, bool some, UT x) {
        if (some) {
            index = rate;
            zoh = combs(x);
        } else {
            index--;
        }
        UT v = zoh;
#pragma unroll
        for (int n = 0; n < N; n++) {
            integ[n] += v;
            v = integ[n];
        }
        return v;
    }
};

template <class UT> __device__ __forceinline__ void unpack16(const int4 &q, UT *e);
template <> __device__ __forceinline__ void unpack16<uint32_t>(const int4 &q, uint32_t *e) {
    e[0] = (uint32_t)q.x; e[1] = (uint32_t)q.y; e[2] = (uint32_t)q.z; e[3] = (uint32_t)q.w;
}
template <> __device__ __forceinline__ void unpack16<uint64_t>(const int4 &q, uint64_t *e) {
    e[0] = (uint64_t)(uint32_t)q.x | ((uint64_t)(uint32_t)q.y << 32);
    e[1] = (uint64_t)(uint32_t)q.z | ((uint64_t)(uint32_t)q.w << 32);
}
template <class UT> __device__ __forceinline__ int4 pack16(const UT *e);
template <> __device__ __forceinline__ int4 pack16<uint32_t>(const uint32_t *e) {
    return make_int4((int)e[0], (int)e[1], (int)e[2], (int)e[3]);
}
template <> __device__ __forceinline__ int4 pack16<uint64_t>(const uint64_t *e) {
    return make_int4((int)(uint32_t)e[0], (int)(uint32_t)(e[0] >> 32), (int)(uint32_t)e[1], (int)(uint32_t)(e[1] >> 32));
}

__device__ __forceinline__ size_t cic_fidx(int layout, size_t t, size_t lane, size_t frames, size_t lanes) {
    return layout == IDSP_FRAME_MAJOR ? t * lanes + lane : lane * frames + t;
}

template <class T, int N, int M>
__global__ void __launch_bounds__(128)
cic_dec_kernel(CicParams<T> p, const T *x, T *y, size_t frames, size_t lanes, size_t sstride, int layout, int vec) {
    using UT = typename std::make_unsigned<T>::type;
    constexpr int V = 16 / sizeof(T);
    const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    CicRegs<T, N, M> c;
    c.load(p, lane, sstride);
    const size_t R = (size_t)p.rate + 1;
    for (size_t t = 0; t < frames; t++) {
        const size_t f = cic_fidx(layout, t, lane, frames, lanes);
        const T *px = x + f * R;
        if (vec) {
#pragma unroll 1
            for (size_t j = 0; j < R; j += V) {
                const int4 q = *reinterpret_cast<const int4 *>(px + j);
                UT e[V];
                unpack16<UT>(q, e);
#pragma unroll
                for (int k = 0; k < V; k++) c.dec_step(p.rate, e[k]);
            }
        } else {
#pragma unroll 1
            for (size_t j = 0; j < R; j++) c.dec_step(p.rate, (UT)px[j]);
        }
        y[f] = (T)c.zoh;  // value of the frame's tick == get_decimate()
    }
    c.store(p, lane, sstride);
}

template <class T, int N, int M>
__global__ void __launch_bounds__(128)
cic_int_kernel(CicParams<T> p, const T *x, T *y, size_t frames, size_t lanes, size_t sstride, int layout, int vec) {
    using UT = typename std::make_unsigned<T>::type;
    constexpr int V = 16 / sizeof(T);
    const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    CicRegs<T, N, M> c;
    c.load(p, lane, sstride);
    const size_t R = (size_t)p.rate + 1;
    for (size_t t = 0; t < frames; t++) {
        const size_t f = cic_fidx(layout, t, lane, frames, lanes);
        T *py = y + f * R;
        const UT xin = (UT)x[f];
        if (vec) {
#pragma unroll 1
            for (size_t j = 0; j < R; j += V) {
                UT e[V];
#pragma unroll
                for (int k = 0; k < V; k++) e[k] = c.int_step(p.rate, j == 0 && k == 0, xin);
                *reinterpret_cast<int4 *>(py + j) = pack16<UT>(e);
            }
        } else {
#pragma unroll 1
            for (size_t j = 0; j < R; j++) py[j] = (T)c.int_step(p.rate, j == 0, xin);
        }
    }
    c.store(p, lane, sstride);
}

// ---------------------------------------------------------------- warp-cooperative variants
// A warp owns 32 lanes and moves tiles of 128 bytes per lane (8 pieces of 16 bytes = 8/P frames,
// P = (rate+1)*sizeof(T)/16 in {1,2,4,8}) with fully coalesced 16-byte accesses: frame-major data is
// one contiguous 32 x P-piece block per frame, lane-major data 128 contiguous bytes per lane.  The
// tile is transposed through shared memory with the piece index XOR-swizzled by (lane & 7), so both
// the cooperative side (consecutive pieces) and the per-lane side (thread = lane, same piece) are
// bank-conflict free.  The next tile is prefetched into registers while the current one is consumed.
// Instantiated for the usual comb delay M = 1 (other M use the direct kernels above).
constexpr int CIC_WPB = 4;  // warps per block

template <class T, int P, bool FM>
__device__ __forceinline__ const T *cic_piece_ptr(const T *x, int c, size_t t0, size_t lane0, size_t frames,
                                                   size_t lanes, int &lane, int &q) {
    constexpr int FT = 8 / P;                   // frames per tile
    constexpr int EPP = 16 / (int)sizeof(T);    // elements per piece
    constexpr int R = P * EPP;                  // rate + 1
    int f, j;
    if (FM) {  // memory order of a frame block: [lane][P]
        f = c / (32 * P);
        lane = (c % (32 * P)) / P;
        j = c % P;
    } else {   // 128 contiguous bytes per lane: [lane][FT][P]
        lane = c / 8;
        f = (c % 8) / P;
        j = c % P;
    }
    (void)FT;
    q = f * P + j;
    const size_t frame = t0 + f;
    const size_t fi = FM ? frame * lanes + lane0 + lane : (lane0 + lane) * frames + frame;
    return x + fi * R + (size_t)j * EPP;
}

template <class T, int N, int M, int P, bool FM>
__global__ void __launch_bounds__(CIC_WPB * 32)
cic_dec_coop_kernel(CicParams<T> p, const T *x, T *y, size_t frames, size_t lanes, size_t sstride) {
    using UT = typename std::make_unsigned<T>::type;
    constexpr int FT = 8 / P, EPP = 16 / (int)sizeof(T);
    __shared__ int4 tile[CIC_WPB][32 * 8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const size_t lane0 = ((size_t)blockIdx.x * CIC_WPB + w) * 32;
    if (lane0 >= lanes) return;
    const int nl = (int)((lanes - lane0) < 32 ? (lanes - lane0) : 32);
    const bool active = l < nl;
    CicRegs<T, N, M> c;
    if (active) c.load(p, lane0 + l, sstride);
    const size_t ntiles = frames / FT;
    int4 pre[8];
    auto fetch = [&](size_t tile_i) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int lane, q;
            const T *src = cic_piece_ptr<T, P, FM>(x, l + 32 * k, tile_i * FT, lane0, frames, lanes, lane, q);
            pre[k] = lane < nl ? *reinterpret_cast<const int4 *>(src) : make_int4(0, 0, 0, 0);
        }
    };
    if (ntiles) fetch(0);
    for (size_t ti = 0; ti < ntiles; ti++) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int lane, q;
            cic_piece_ptr<T, P, FM>(x, l + 32 * k, 0, 0, frames, lanes, lane, q);
            tile[w][lane * 8 + (q ^ (lane & 7))] = pre[k];
        }
        __syncwarp();
        if (ti + 1 < ntiles) fetch(ti + 1);
        if (active) {
#pragma unroll
            for (int f = 0; f < FT; f++) {
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const int4 v = tile[w][l * 8 + ((f * P + j) ^ (l & 7))];
                    UT e[EPP];
                    unpack16<UT>(v, e);
#pragma unroll
                    for (int k = 0; k < EPP; k++) c.dec_step(p.rate, e[k]);
                }
                const size_t frame = ti * FT + f;
                y[FM ? frame * lanes + lane0 + l : (lane0 + l) * frames + frame] = (T)c.zoh;
            }
        }
    }
    // frames that do not fill a tile
    if (active) {
        constexpr int R = P * EPP;
        for (size_t frame = ntiles * FT; frame < frames; frame++) {
            const size_t fi = FM ? frame * lanes + lane0 + l : (lane0 + l) * frames + frame;
            for (int j = 0; j < R; j++) c.dec_step(p.rate, (UT)x[fi * R + j]);
            y[fi] = (T)c.zoh;
        }
        c.store(p, lane0 + l, sstride);
    }
}

template <class T, int N, int M, int P, bool FM>
__global__ void __launch_bounds__(CIC_WPB * 32)
cic_int_coop_kernel(CicParams<T> p, const T *x, T *y, size_t frames, size_t lanes, size_t sstride) {
    using UT = typename std::make_unsigned<T>::type;
    constexpr int FT = 8 / P, EPP = 16 / (int)sizeof(T);
    __shared__ int4 tile[CIC_WPB][32 * 8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const size_t lane0 = ((size_t)blockIdx.x * CIC_WPB + w) * 32;
    if (lane0 >= lanes) return;
    const int nl = (int)((lanes - lane0) < 32 ? (lanes - lane0) : 32);
    const bool active = l < nl;
    CicRegs<T, N, M> c;
    if (active) c.load(p, lane0 + l, sstride);
    const size_t ntiles = frames / FT;
    for (size_t ti = 0; ti < ntiles; ti++) {
        __syncwarp();
        if (active) {
#pragma unroll
            for (int f = 0; f < FT; f++) {
                const size_t frame = ti * FT + f;
                const UT xin = (UT)x[FM ? frame * lanes + lane0 + l : (lane0 + l) * frames + frame];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    UT e[EPP];
#pragma unroll
                    for (int k = 0; k < EPP; k++) e[k] = c.int_step(p.rate, j == 0 && k == 0, xin);
                    tile[w][l * 8 + ((f * P + j) ^ (l & 7))] = pack16<UT>(e);
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int lane, q;
            const T *dst = cic_piece_ptr<T, P, FM>(y, l + 32 * k, ti * FT, lane0, frames, lanes, lane, q);
            if (lane < nl) *reinterpret_cast<int4 *>(const_cast<T *>(dst)) = tile[w][lane * 8 + (q ^ (lane & 7))];
        }
    }
    if (active) {
        constexpr int R = P * EPP;
        for (size_t frame = ntiles * FT; frame < frames; frame++) {
            const size_t fi = FM ? frame * lanes + lane0 + l : (lane0 + l) * frames + frame;
            const UT xin = (UT)x[fi];
            for (int j = 0; j < R; j++) y[fi * R + j] = (T)c.int_step(p.rate, j == 0, xin);
        }
        c.store(p, lane0 + l, sstride);
    }
}

template <class T, bool DEC, int N, int M>
static bool cic_launch_coop(idsp_ctx *ctx, const CicParams<T> &p, const T *x, T *y, size_t frames, size_t lanes,
                            int layout) {
    const size_t bytes = ((size_t)p.rate + 1) * sizeof(T);
    if (bytes != 16 && bytes != 32 && bytes != 64 && bytes != 128) return false;
    const unsigned grid = (unsigned)((lanes + 32 * CIC_WPB - 1) / (32 * CIC_WPB));
    const bool fm = layout == IDSP_FRAME_MAJOR;
#define COOP(PP)                                                                                                    \
    do {                                                                                                            \
        if (DEC) {                                                                                                  \
            if (fm) cic_dec_coop_kernel<T, N, M, PP, true><<<grid, 32 * CIC_WPB, 0, ctx->stream>>>(p, x, y, frames, lanes, lanes); \
            else cic_dec_coop_kernel<T, N, M, PP, false><<<grid, 32 * CIC_WPB, 0, ctx->stream>>>(p, x, y, frames, lanes, lanes);   \
        } else {                                                                                                    \
            if (fm) cic_int_coop_kernel<T, N, M, PP, true><<<grid, 32 * CIC_WPB, 0, ctx->stream>>>(p, x, y, frames, lanes, lanes); \
            else cic_int_coop_kernel<T, N, M, PP, false><<<grid, 32 * CIC_WPB, 0, ctx->stream>>>(p, x, y, frames, lanes, lanes);   \
        }                                                                                                           \
    } while (0)
    switch (bytes) {
        case 16: COOP(1); break;
        case 32: COOP(2); break;
        case 64: COOP(4); break;
        default: COOP(8); break;
    }
#undef COOP
    IDSP_KERNEL_FAMILY(ctx, "cic warp-cooperative tiles");
    return true;
}

template <class T, bool DEC>
int cic_launch(idsp_ctx *ctx, int N, int M, uint32_t rate, T *state, const T *x, T *y, size_t frames, size_t lanes,
               int layout) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    IDSP_CHECK_ARG(N >= 1 && N <= 6, "order N must be 1..6");
    IDSP_CHECK_ARG(M >= 1 && M <= CIC_MAXM, "comb delay M must be 1..3");
    IDSP_CHECK_ARG(layout == IDSP_FRAME_MAJOR || layout == IDSP_LANE_MAJOR, "layout must be 0 (frame-major) or 1 (lane-major)");
    if (frames == 0 || lanes == 0) return IDSP_OK;
    IDSP_CHECK_ARG(state && x && y, "state/x/y must not be null");
    CicParams<T> p{state, rate, M};
    const size_t R = (size_t)rate + 1;
    const T *wide = DEC ? x : y;  // the rate+1 wide side
    const int vec = ((R * sizeof(T)) % 16 == 0 && (((uintptr_t)wide) & 15) == 0) ? 1 : 0;
    const unsigned grid = (unsigned)((lanes + 127) / 128);
    IDSP_KERNEL_FAMILY(ctx, "cic thread-per-lane");
#define GO(NN, MM)                                                                                                  \
    case NN * 8 + MM:                                                                                               \
        if (MM == 1 && vec && ctx->policy != 1 && cic_launch_coop<T, DEC, NN, 1>(ctx, p, x, y, frames, lanes, layout)) break; \
        if (DEC) cic_dec_kernel<T, NN, MM><<<grid, 128, 0, ctx->stream>>>(p, x, y, frames, lanes, lanes, layout, vec); \
        else cic_int_kernel<T, NN, MM><<<grid, 128, 0, ctx->stream>>>(p, x, y, frames, lanes, lanes, layout, vec);    \
        break;
#define GOM(NN) GO(NN, 1) GO(NN, 2) GO(NN, 3)
    switch (N * 8 + M) {
        GOM(1) GOM(2) GOM(3) GOM(4) GOM(5) GOM(6)
    }
#undef GOM
#undef GO
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

}  // namespace

extern "C" size_t idsp_cic_state_words(int N, int M) { return (size_t)(2 + N * M + N); }
extern "C" int idsp_cic_dec_i32(idsp_ctx *ctx, int N, int M, uint32_t rate, int32_t *state, const int32_t *x,
                                int32_t *y, size_t frames, size_t lanes, int layout) {
    return cic_launch<int32_t, true>(ctx, N, M, rate, state, x, y, frames, lanes, layout);
}
extern "C" int idsp_cic_dec_i64(idsp_ctx *ctx, int N, int M, uint32_t rate, int64_t *state, const int64_t *x,
                                int64_t *y, size_t frames, size_t lanes, int layout) {
    return cic_launch<int64_t, true>(ctx, N, M, rate, state, x, y, frames, lanes, layout);
}
extern "C" int idsp_cic_int_i32(idsp_ctx *ctx, int N, int M, uint32_t rate, int32_t *state, const int32_t *x,
                                int32_t *y, size_t frames, size_t lanes, int layout) {
    return cic_launch<int32_t, false>(ctx, N, M, rate, state, x, y, frames, lanes, layout);
}
extern "C" int idsp_cic_int_i64(idsp_ctx *ctx, int N, int M, uint32_t rate, int64_t *state, const int64_t *x,
                                int64_t *y, size_t frames, size_t lanes, int layout) {
    return cic_launch<int64_t, false>(ctx, N, M, rate, state, x, y, frames, lanes, layout);
}
