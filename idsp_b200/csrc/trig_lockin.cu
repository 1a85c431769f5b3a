// trig_lockin.cu -- cossin / atan2 / Lowpass / Lockin entry points (include/idsp_b200.h).
#include "common.cuh"
#include "ops.cuh"
#include "lane_kernels.cuh"
#include "tma_kernels.cuh"

using namespace idsp;

// ---------------------------------------------------------------- memoryless maps
// 2 phases per thread: one 8-byte load and ONE 16-byte store, so a warp store covers 512 contiguous
// bytes (with 4 phases per thread the two 16-byte stores of a thread interleave with its neighbours'
// and every 32-byte sector is written in two halves); two such pairs in flight per thread.
__global__ void __launch_bounds__(256) cossin_kernel(const int32_t *phase, int32_t *cs, size_t n) {
    __shared__ __align__(8) uint32_t lut_all[256 * IDSP_COSSIN_REP];
    cossin_expand_lut<IDSP_COSSIN_REP>(g_cossin_lut, lut_all, threadIdx.x, blockDim.x);
    __syncthreads();
    const uint32_t *lut = lut_all + 2 * (threadIdx.x % IDSP_COSSIN_REP);  // this thread's copy (conflict-free lookups)
    const size_t n2 = n / 2;
    const bool vec = ((((uintptr_t)phase) & 7) | (((uintptr_t)cs) & 15)) == 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        // software pipeline: the loads of the next pair are issued before the current pair is converted
        int2 p = make_int2(0, 0), q = make_int2(0, 0);
        if (i + stride < n2) {
            p = reinterpret_cast<const int2 *>(phase)[i];
            q = reinterpret_cast<const int2 *>(phase)[i + stride];
        }
        for (; i + stride < n2; i += 2 * stride) {
            int2 pn = p, qn = q;
            if (i + 3 * stride < n2) {
                pn = reinterpret_cast<const int2 *>(phase)[i + 2 * stride];
                qn = reinterpret_cast<const int2 *>(phase)[i + 3 * stride];
            }
            int4 a, b;
            cossin_dev_x<IDSP_COSSIN_REP>(lut, p.x, a.x, a.y);
            cossin_dev_x<IDSP_COSSIN_REP>(lut, p.y, a.z, a.w);
            cossin_dev_x<IDSP_COSSIN_REP>(lut, q.x, b.x, b.y);
            cossin_dev_x<IDSP_COSSIN_REP>(lut, q.y, b.z, b.w);
            reinterpret_cast<int4 *>(cs)[i] = a;
            reinterpret_cast<int4 *>(cs)[i + stride] = b;
            p = pn;
            q = qn;
        }
        for (; i < n2; i += stride) {
            const int2 p = reinterpret_cast<const int2 *>(phase)[i];
            int4 a;
            cossin_dev_x<IDSP_COSSIN_REP>(lut, p.x, a.x, a.y);
            cossin_dev_x<IDSP_COSSIN_REP>(lut, p.y, a.z, a.w);
            reinterpret_cast<int4 *>(cs)[i] = a;
        }
        i = n2 * 2 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    }
    for (; i < n; i += stride) {
        int32_t c, s;
        cossin_dev_x<IDSP_COSSIN_REP>(lut, phase[i], c, s);
        cs[2 * i] = c;
        cs[2 * i + 1] = s;
    }
}
__global__ void __launch_bounds__(256) atan2_kernel(const int32_t *xy, int32_t *p, size_t n) {
    __shared__ uint2 tab[16];  // reciprocal seeds staged per CTA (LDS.64 instead of a global load per call)
    if (threadIdx.x < 16) tab[threadIdx.x] = g_divi_tab[threadIdx.x];
    __syncthreads();
    auto atan2_dev = [&](int32_t y, int32_t x) { return atan2_dev_t<true>(tab, y, x); };
    const size_t n2 = n / 2;
    const bool vec = ((((uintptr_t)xy) & 15) | (((uintptr_t)p) & 7)) == 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        // two 16-byte loads in flight per thread, software pipelined: the next pair is requested before the
        // current one is converted (the conversion is ~240 instructions long, and without this the warps
        // sit on the long scoreboard at the top of every iteration)
        int4 v0 = make_int4(0, 0, 0, 0), v1 = v0;
        if (i + stride < n2) {
            v0 = reinterpret_cast<const int4 *>(xy)[i];
            v1 = reinterpret_cast<const int4 *>(xy)[i + stride];
        }
        for (; i + stride < n2; i += 2 * stride) {
            int4 n0 = v0, n1 = v1;
            if (i + 3 * stride < n2) {
                n0 = reinterpret_cast<const int4 *>(xy)[i + 2 * stride];
                n1 = reinterpret_cast<const int4 *>(xy)[i + 3 * stride];
            }
            int2 r0, r1;
            r0.x = atan2_dev(v0.y, v0.x);
            r0.y = atan2_dev(v0.w, v0.z);
            r1.x = atan2_dev(v1.y, v1.x);
            r1.y = atan2_dev(v1.w, v1.z);
            reinterpret_cast<int2 *>(p)[i] = r0;
            reinterpret_cast<int2 *>(p)[i + stride] = r1;
            v0 = n0;
            v1 = n1;
        }
        for (; i < n2; i += stride) {
            int4 v = reinterpret_cast<const int4 *>(xy)[i];
            int2 r;
            r.x = atan2_dev(v.y, v.x);
            r.y = atan2_dev(v.w, v.z);
            reinterpret_cast<int2 *>(p)[i] = r;
        }
        i = n2 * 2 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    }
    for (; i < n; i += stride) p[i] = atan2_dev(xy[2 * i + 1], xy[2 * i]);
}

static unsigned map_grid(idsp_ctx *ctx, size_t work_items) {
    size_t blocks = (work_items + 255) / 256;
    size_t cap = (size_t)ctx->sm_count * 8;  // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

extern "C" int idsp_cossin_i32(idsp_ctx *ctx, const int32_t *phase, int32_t *cs, size_t n) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    if (n == 0) return IDSP_OK;
    IDSP_CHECK_ARG(phase && cs, "phase/cs must not be null");
    cossin_kernel<<<map_grid(ctx, (n + 1) / 2), 256, 0, ctx->stream>>>(phase, cs, n);
    IDSP_KERNEL_FAMILY(ctx, "map cossin");
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}
extern "C" int idsp_atan2_i32(idsp_ctx *ctx, const int32_t *xy, int32_t *p, size_t n) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    if (n == 0) return IDSP_OK;
    IDSP_CHECK_ARG(xy && p, "xy/p must not be null");
    atan2_kernel<<<map_grid(ctx, (n + 1) / 2), 256, 0, ctx->stream>>>(xy, p, n);
    IDSP_KERNEL_FAMILY(ctx, "map atan2");
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}
extern "C" int idsp_cossin_i32_host(idsp_ctx *ctx, const int32_t *phase, int32_t *cs, size_t n) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    if (n == 0) return IDSP_OK;
    IDSP_CHECK_ARG(phase && cs, "phase/cs must not be null");
    HostStreamSpec s;
    s.frames = n;
    s.lanes = 1;
    s.in_bytes_per_frame_lane = 4;
    s.out_bytes_per_frame_lane = 8;
    s.layout = IDSP_FRAME_MAJOR;
    s.nblobs = 0;
    return idsp_host_stream(ctx, s, phase, cs,
                            [&](void **, const void *dx, void *dy, size_t, size_t an) {
                                return idsp_cossin_i32(ctx, (const int32_t *)dx, (int32_t *)dy, an);
                            });
}
extern "C" int idsp_atan2_i32_host(idsp_ctx *ctx, const int32_t *xy, int32_t *p, size_t n) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    if (n == 0) return IDSP_OK;
    IDSP_CHECK_ARG(xy && p, "xy/p must not be null");
    HostStreamSpec s;
    s.frames = n;
    s.lanes = 1;
    s.in_bytes_per_frame_lane = 8;
    s.out_bytes_per_frame_lane = 4;
    s.layout = IDSP_FRAME_MAJOR;
    s.nblobs = 0;
    return idsp_host_stream(ctx, s, xy, p,
                            [&](void **, const void *dx, void *dy, size_t, size_t an) {
                                return idsp_atan2_i32(ctx, (const int32_t *)dx, (int32_t *)dy, an);
                            });
}

// ---------------------------------------------------------------- Lowpass / Lockin
#define LL_CHECK()                                                                   \
    do {                                                                             \
        int r_ = idsp_use_device(ctx);                                               \
        if (r_) return r_;                                                           \
        IDSP_CHECK_ARG(order == 1 || order == 2, "order must be 1 or 2 (lowpass.rs:74-76)"); \
        IDSP_CHECK_ARG(k != nullptr, "k is null");                                   \
        IDSP_CHECK_ARG(layout == IDSP_FRAME_MAJOR || layout == IDSP_LANE_MAJOR,      \
                       "layout must be 0 (frame-major) or 1 (lane-major)");          \
        if (frames == 0 || lanes == 0) return IDSP_OK;                               \
    } while (0)

extern "C" int idsp_lowpass_i32(idsp_ctx *ctx, int order, const int32_t *k, int64_t *state,
                                const int32_t *x, int32_t *y, size_t frames, size_t lanes,
                                int layout) {
    LL_CHECK();
    IDSP_CHECK_ARG(state && x && y, "state/x/y must not be null");
    if (order == 1) {
        LowpassOp<1>::Params p;
        p.k[0] = p.kk[0] = k[0];
        p.k[1] = p.kk[1] = 0;
        p.st = state;
        return launch_lanes_best<LowpassOp<1>>(ctx, p, x, y, frames, lanes, lanes, layout);
    }
    LowpassOp<2>::Params p;
    p.k[0] = p.kk[0] = k[0];
    p.k[1] = p.kk[1] = k[1];
    p.st = state;
    return launch_lanes_best<LowpassOp<2>>(ctx, p, x, y, frames, lanes, lanes, layout);
}

int lockin_dev(idsp_ctx *ctx, int order, const int32_t *k, int32_t *accu_state,
               const int32_t *accu_step, int64_t *lp_state, const int32_t *x, int32_t *iq,
               size_t frames, size_t lanes, size_t sstride, int layout) {
    const uint32_t *lut;
    IDSP_CUDA(cudaGetSymbolAddress((void **)&lut, g_cossin_lut));
#define GO(ORDER)                                                                    \
    do {                                                                             \
        LockinOp<ORDER, true>::Params pt;                                            \
        pt.k[0] = pt.kk[0] = k[0];                                                   \
        pt.k[1] = pt.kk[1] = ORDER == 2 ? k[1] : 0;                                  \
        pt.accu_state = accu_state;                                                  \
        pt.accu_step = accu_step;                                                    \
        pt.st = lp_state;                                                            \
        pt.lut = lut;                                                                \
        int tr = tma_try_launch<LockinOp<ORDER, true>>(ctx, pt, x, (int2 *)iq, frames, lanes, sstride, layout); \
        if (tr != IDSP_TMA_NOT_APPLICABLE) return tr;                                \
        LockinOp<ORDER, false>::Params pg;                                           \
        pg.k[0] = pg.kk[0] = pt.k[0];                                                \
        pg.k[1] = pg.kk[1] = pt.k[1];                                                \
        pg.accu_state = accu_state;                                                  \
        pg.accu_step = accu_step;                                                    \
        pg.st = lp_state;                                                            \
        pg.lut = lut;                                                                \
        return launch_lanes<LockinOp<ORDER, false>>(ctx, pg, x, (int2 *)iq, frames, lanes, sstride, layout); \
    } while (0)
    if (order == 1) GO(1);
    GO(2);
#undef GO
}

extern "C" int idsp_lockin_i32(idsp_ctx *ctx, int order, const int32_t *k, int32_t *accu_state,
                               const int32_t *accu_step, int64_t *lp_state, const int32_t *x,
                               int32_t *iq, size_t frames, size_t lanes, int layout) {
    LL_CHECK();
    IDSP_CHECK_ARG(accu_state && accu_step && lp_state && x && iq, "null pointer argument");
    IDSP_CHECK_ARG((((uintptr_t)iq) & 7) == 0, "iq must be 8-byte aligned");
    return lockin_dev(ctx, order, k, accu_state, accu_step, lp_state, x, iq, frames, lanes, lanes,
                      layout);
}

extern "C" int idsp_lockin_i32_host(idsp_ctx *ctx, int order, const int32_t *k,
                                    int32_t *accu_state, const int32_t *accu_step,
                                    int64_t *lp_state, const int32_t *x, int32_t *iq, size_t frames,
                                    size_t lanes, int layout) {
    LL_CHECK();
    IDSP_CHECK_ARG(accu_state && accu_step && lp_state && x && iq, "null pointer argument");
    HostStreamSpec s;
    s.frames = frames;
    s.lanes = lanes;
    s.in_bytes_per_frame_lane = 4;
    s.out_bytes_per_frame_lane = 8;
    s.layout = layout;
    s.nblobs = 3;
    s.blobs[0] = {accu_state, lanes * 4, true};
    s.blobs[1] = {const_cast<int32_t *>(accu_step), lanes * 4, false};
    s.blobs[2] = {lp_state, (size_t)2 * order * lanes * 8, true};
    return idsp_host_stream(
        ctx, s, x, iq, [&](void **b, const void *dx, void *dy, size_t a0, size_t an) {
            if (layout == IDSP_FRAME_MAJOR)
                return lockin_dev(ctx, order, k, (int32_t *)b[0], (const int32_t *)b[1],
                                  (int64_t *)b[2], (const int32_t *)dx, (int32_t *)dy, an, lanes,
                                  lanes, layout);
            return lockin_dev(ctx, order, k, (int32_t *)b[0] + a0, (const int32_t *)b[1] + a0,
                              (int64_t *)b[2] + a0, (const int32_t *)dx, (int32_t *)dy, frames, an,
                              lanes, layout);
        });
}

// (sample, phase) tuples: src/lockin.rs:30-39
extern "C" int idsp_lockin_phase_i32(idsp_ctx *ctx, int order, const int32_t *k, int64_t *lp_state,
                                     const int32_t *xp, int32_t *iq, size_t frames, size_t lanes, int layout) {
    LL_CHECK();
    IDSP_CHECK_ARG(lp_state && xp && iq, "null pointer argument");
    IDSP_CHECK_ARG(((((uintptr_t)xp) | ((uintptr_t)iq)) & 7) == 0, "xp and iq must be 8-byte aligned");
    const uint32_t *lut;
    IDSP_CUDA(cudaGetSymbolAddress((void **)&lut, g_cossin_lut));
#define GO(ORDER)                                                                    \
    do {                                                                             \
        LockinPhaseOp<ORDER, true>::Params pt;                                       \
        pt.k[0] = pt.kk[0] = k[0];                                                   \
        pt.k[1] = pt.kk[1] = ORDER == 2 ? k[1] : 0;                                  \
        pt.st = lp_state;                                                            \
        pt.lut = lut;                                                                \
        int tr = tma_try_launch<LockinPhaseOp<ORDER, true>>(ctx, pt, (const int2 *)xp, (int2 *)iq, frames, lanes, lanes, layout); \
        if (tr != IDSP_TMA_NOT_APPLICABLE) return tr;                                \
        LockinPhaseOp<ORDER, false>::Params pg;                                      \
        pg.k[0] = pg.kk[0] = pt.k[0];                                                \
        pg.k[1] = pg.kk[1] = pt.k[1];                                                \
        pg.st = lp_state;                                                            \
        pg.lut = lut;                                                                \
        return launch_lanes<LockinPhaseOp<ORDER, false>>(ctx, pg, (const int2 *)xp, (int2 *)iq, frames, lanes, lanes, layout); \
    } while (0)
    if (order == 1) GO(1);
    GO(2);
#undef GO
}

// (sample, LO) tuples: src/lockin.rs:17-28
extern "C" int idsp_lockin_lo_i32(idsp_ctx *ctx, int order, const int32_t *k, int64_t *lp_state,
                                  const int32_t *xlo, int32_t *iq, size_t frames, size_t lanes, int layout) {
    LL_CHECK();
    IDSP_CHECK_ARG(lp_state && xlo && iq, "null pointer argument");
    IDSP_CHECK_ARG((((uintptr_t)iq) & 7) == 0, "iq must be 8-byte aligned");
    if (order == 1) {
        LockinLoOp<1>::Params p;
        p.k[0] = p.kk[0] = k[0];
        p.k[1] = p.kk[1] = 0;
        p.st = lp_state;
        return launch_lanes<LockinLoOp<1>>(ctx, p, (const XLo *)xlo, (int2 *)iq, frames, lanes, lanes, layout);
    }
    LockinLoOp<2>::Params p;
    p.k[0] = p.kk[0] = k[0];
    p.k[1] = p.kk[1] = k[1];
    p.st = lp_state;
    return launch_lanes<LockinLoOp<2>>(ctx, p, (const XLo *)xlo, (int2 *)iq, frames, lanes, lanes, layout);
}

// ---------------------------------------------------------------- PLL
extern "C" int idsp_pll_i32(idsp_ctx *ctx, const int32_t *ba, int32_t *state, const int32_t *x, int32_t *y,
                            size_t frames, size_t lanes, int layout) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    IDSP_CHECK_ARG(ba != nullptr, "ba is null");
    IDSP_CHECK_ARG(layout == IDSP_FRAME_MAJOR || layout == IDSP_LANE_MAJOR,
                   "layout must be 0 (frame-major) or 1 (lane-major)");
    if (frames == 0 || lanes == 0) return IDSP_OK;
    IDSP_CHECK_ARG(state && x && y, "state/x/y must not be null");
    PllOp::Params p;
    p.ba[0] = ba[0];
    p.ba[1] = ba[1];
    p.ba[2] = ba[2];
    p.st = state;
    return launch_lanes_best<PllOp>(ctx, p, x, y, frames, lanes, lanes, layout);
}

// ---------------------------------------------------------------- FM discriminator
extern "C" int idsp_fm_disc_i32(idsp_ctx *ctx, int32_t carrier, const int32_t *ba, int F, int32_t *state,
                                const int32_t *x, int32_t *y, size_t frames, size_t lanes, int layout) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    IDSP_CHECK_ARG(ba != nullptr, "ba is null");
    IDSP_CHECK_ARG(F > -32 && F < 64, "F out of range");
    IDSP_CHECK_ARG(layout == IDSP_FRAME_MAJOR || layout == IDSP_LANE_MAJOR,
                   "layout must be 0 (frame-major) or 1 (lane-major)");
    if (frames == 0 || lanes == 0) return IDSP_OK;
    IDSP_CHECK_ARG(state && x && y, "state/x/y must not be null");
    IDSP_CHECK_ARG((((uintptr_t)x) & 7) == 0, "x (re, im pairs) must be 8-byte aligned");
    if (F >= 0 && F < 32) {
        FmDiscOp<1>::Params p;
        p.carrier = carrier;
        for (int i = 0; i < 5; i++) p.ba[i] = ba[i];
        p.F = F;
        p.st = state;
        return launch_lanes_best<FmDiscOp<1>>(ctx, p, (const int2 *)x, y, frames, lanes, lanes, layout);
    }
    FmDiscOp<0>::Params p;
    p.carrier = carrier;
    for (int i = 0; i < 5; i++) p.ba[i] = ba[i];
    p.F = F;
    p.st = state;
    return launch_lanes_best<FmDiscOp<0>>(ctx, p, (const int2 *)x, y, frames, lanes, lanes, layout);
}
