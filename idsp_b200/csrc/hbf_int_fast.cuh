// hbf_int_fast.cuh -- shared-memory tiled HBF x2^K interpolation cascade (f32, lane-major).
//
// Mirror image of hbf_fast_scalar.cuh.  A CTA owns NL = 8 lanes and walks time in tiles of
// TI = 512 >> K input samples per lane (512 output samples per lane per tile):
//
//   x tile (128 B per lane) --register prefetch--> X rows [hist | TI]
//   stage 0 (TAPS[0], lowest rate) : X -> U_1           (shared memory -> shared memory)
//   stage s                        : U_s -> U_{s+1}
//   stage K-1 (highest rate)       : U_{K-1} -> output staging ring (2 x 2 KB per lane)
//   staging --TMA 1-D bulk store per lane row (2 KB, bulk_group)--> HBM
//
// Work item = (lane, R consecutive inputs of one stage) -> 2R outputs:
//   out[2n]   = ((x[n]+x[n-2M+1])*c0) + ((x[n-1]+x[n-2M+2])*c1) + ...   (src/hbf.rs:46-68, :220-224)
//   out[2n+1] = x[n-M+1]                                              (centre tap = identity)
// each op individually rounded, summed in the reference's order -> bit-exact.  Items are
// spread lane-fastest over the threads; rows have a pitch of 4*odd floats so every
// LDS.128 / STS.128 over 8 lanes is conflict free.  Rows are [history | tile]; tails move to
// heads after each tile (copy_within, src/hbf.rs:231); ABI state <-> those heads at entry/exit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "hbf_stages.cuh"
#include "ops.cuh"
#include "tma_kernels.cuh"

#ifndef HFI_BQ_MINB
#define HFI_BQ_MINB 4
#endif

namespace idsp {
__device__ __forceinline__ float4 lds128v(const float *p) {  // volatile: never narrowed by ptxas
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(smem_u32(p)));
    return v;
}
// named barriers of the fused biquad variant (id 0 = __syncthreads): 1 = the FIR warps among themselves,
// 2 + b = staging buffer b is full (FIR warps arrive, the biquad warp waits), 4 + b = buffer b is free again
__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bulk_store_1d(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}
}  // namespace idsp

#define HFI_TAPS HbfTaps
#define HFI_M hbf_m
// default shape: 8 lanes x 512 outputs per tile
#define HFI_NS hfi
#define HFI_NL 8
#define HFI_TOUT 512
#include "hbf_int_fast_body.cuh"
#undef HFI_NS
#undef HFI_NL
#undef HFI_TOUT
// fused-biquad shape: 16 lanes x 256 outputs per tile (same shared memory): the biquad warp advances all
// the lanes of a CTA with every instruction, so twice the lanes halve its share of the tile time
#define HFI_NS hfi16
#define HFI_NL 16
#define HFI_TOUT 256
#include "hbf_int_fast_body.cuh"
#undef HFI_NS
#undef HFI_NL
#undef HFI_TOUT
#ifdef IDSP_TUNE
// tuning builds: 32 lanes x 256 outputs per tile, 8 FIR warps, two CTAs per SM (IDSP_CHAIN_WIDE=2)
#define HFI_NS hfi32
#define HFI_NL 32
#define HFI_TOUT 256
#define HFI_NT 256
#define HFI_MINB 2
#include "hbf_int_fast_body.cuh"
#undef HFI_NS
#undef HFI_NL
#undef HFI_TOUT
#undef HFI_NT
#undef HFI_MINB
#endif
#undef HFI_TAPS
#undef HFI_M
// HBF_TAPS_98 (src/hbf.rs:258-292) on the default shape
#define HFI_TAPS HbfTaps98
#define HFI_M hbf98_m
#define HFI_NS hfi98
#define HFI_NL 8
#define HFI_TOUT 512
#include "hbf_int_fast_body.cuh"
#undef HFI_NS
#undef HFI_NL
#undef HFI_TOUT
#undef HFI_TAPS
#undef HFI_M

namespace idsp {
// Runs the tiled kernel over the first (n_in / TI) * TI input frames of every lane; *done =
// frames covered (the generic kernel finishes the tail), or IDSP_HBF_FAST_NOT_APPLICABLE.
// (frame-major x32 with per-thread 16-byte stores measured slower than the generic thread-per-lane kernel, 560 vs
// 714 GSa/s; with tensor-map output tiles the tiled kernel is the default for x32 too)
#define IDSP_DEF_INT_FAST_TRY(NAME, NS) \
    static int NAME(idsp_ctx *ctx, int k, float *state, const float *x, float *y, size_t n_in, \
                                size_t lanes, size_t sstride, int layout, size_t *done) { \
        *done = 0; \
        const bool fm = layout == IDSP_FRAME_MAJOR; \
        if (ctx->policy == 1 || (fm && k > 4 && ctx->policy != 2 && getenv("IDSP_HBF_FM_LDGSTS"))) return IDSP_HBF_FAST_NOT_APPLICABLE; \
        const size_t TI = (size_t)NS::TOUT >> k; \
        const size_t ntiles = n_in / TI; \
        const bool ok = ntiles >= 1 && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0 && (fm || (n_in % 4) == 0); \
        if (!ok) return IDSP_HBF_FAST_NOT_APPLICABLE; \
        int r; \
        if (fm) { \
            switch (k) { \
                case 1: r = NS::launch<1, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
                case 2: r = NS::launch<2, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
                case 3: r = NS::launch<3, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
                case 4: r = NS::launch<4, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
                default: r = NS::launch<5, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
            } \
        } else { \
            switch (k) { \
                case 1: r = NS::launch<1, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
                case 2: r = NS::launch<2, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
                case 3: r = NS::launch<3, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
                case 4: r = NS::launch<4, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
                default: r = NS::launch<5, false>(ctx, state, x, y, n_in, ntiles, lanes, sstride); break; \
            } \
        } \
        if (r == IDSP_OK) *done = ntiles * TI; \
        return r; \
    }
IDSP_DEF_INT_FAST_TRY(hbf_int_fast_try, hfi)
IDSP_DEF_INT_FAST_TRY(hbf98_int_fast_try, hfi98)
#undef IDSP_DEF_INT_FAST_TRY

// HbfInt x2^k -> Biquad DF1 f32 in one pass (lane-major, whole tiles only); `bq.st` = the biquad's SoA
// state [x1, x2, y1, y2][sstride].  wide = false: 8 lanes x 512 outputs per CTA tile (most CTAs: few
// lanes), wide = true: 16 x 256 (the biquad warp advances 16 lanes per instruction: many lanes).
static size_t hbf_int_bq_tile(int k, int wide) { return (size_t)(wide ? hfi16::TOUT : hfi::TOUT) >> k; }
static int hbf_int_bq_fast_try(idsp_ctx *ctx, int k, float *state, const float *x, float *y, size_t n_in, size_t lanes,
                               size_t sstride, const Df1Op<float, false>::Params &bq, int wide) {
    const size_t TI = hbf_int_bq_tile(k, wide);
    if (ctx->policy == 1 || n_in == 0 || n_in % TI != 0 || ((((uintptr_t)x) | ((uintptr_t)y)) & 15) != 0)
        return IDSP_HBF_FAST_NOT_APPLICABLE;
    const size_t ntiles = n_in / TI;
#define IDSP_GO(NS)                                                                                       \
    switch (k) {                                                                                          \
        case 1: return NS::launch<1, false, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride, bq);    \
        case 2: return NS::launch<2, false, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride, bq);    \
        case 3: return NS::launch<3, false, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride, bq);    \
        case 4: return NS::launch<4, false, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride, bq);    \
        default: return NS::launch<5, false, true>(ctx, state, x, y, n_in, ntiles, lanes, sstride, bq);   \
    }
#ifdef IDSP_TUNE
    if (wide == 2) { IDSP_GO(hfi32) }
#endif
    if (wide) { IDSP_GO(hfi16) }
    IDSP_GO(hfi)
#undef IDSP_GO
}

}  // namespace idsp
