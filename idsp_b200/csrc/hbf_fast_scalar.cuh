// hbf_fast_scalar.cuh -- shared-memory tiled HBF /2^K decimation cascade (f32, lane-major),
// scalar FP32 variant (the default; hbf_fast.cuh holds the packed f32x2 variant, both bit-exact).
//
// The FIR stages are time-parallel, so unlike the biquad kernels a lane is not tied to one
// thread.  A CTA owns NL = 8 lanes for the whole call and walks the time axis in tiles of
// TT = 512 input samples per lane (4 CTAs of 128 threads per SM, each its own barrier domain):
//
//   HBM --TMA 1-D bulk copy per lane row (2 KB), mbarrier complete_tx--> raw ring (S = 2)
//   stage 0 : reads the interleaved raw stream, writes de-interleaved E/O rows of stage 1
//   stage s : reads E_s / O_s, writes E_{s+1} / O_{s+1}          (all in shared memory)
//   stage K-1 : writes the decimated output straight to HBM
//
// Work item = (lane, R consecutive outputs of one stage); items are spread over the CTA's
// threads lane-fastest, so a quarter-warp touches 8 different rows whose pitch is 4*odd
// floats -> every LDS.128 / STS.128 is bank-conflict free.  A thread loads its whole window
// into registers with static indices (no shifting delay line) and evaluates R outputs in
// exactly the reference's order:
//   acc = ((w[2M-1]+w[0])*c0) + ((w[2M-2]+w[1])*c1) + ...  then  + even sample
// (src/hbf.rs:46-68, :178-181), each op individually rounded (-fmad=false) -> bit-exact.
// Each row keeps the history the next tile needs ([hist | tile]); after a tile the tails are
// moved to the heads (the reference's copy_within, src/hbf.rs:183-184).  The ABI state
// (even/odd history per stage) is scattered into those heads at entry and gathered back at
// exit, so calls can be chained like block() calls on the reference.
// Tile-shape measurements (262144 lanes x 65536 inputs, /16): NL16/NT256/R0=8 826 GSa/s,
// NL16/NT256/R0=16 846, NL8/NT64 873, NL8/NT128/R0=16 901 (profiles/r1_hbf_variants.log).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "hbf_stages.cuh"
#include "tma_kernels.cuh"

#ifndef IDSP_HBF_FAST_NOT_APPLICABLE
#define IDSP_HBF_FAST_NOT_APPLICABLE 12346
#endif


// HBF_TAPS (src/hbf.rs:308-349): the reference's cascades
#define HFS_NS hfs
#define HFS_TAPS HbfTaps
#define HFS_M hbf_m
#include "hbf_fast_scalar_body.cuh"
#undef HFS_NS
#undef HFS_TAPS
#undef HFS_M
// HBF_TAPS_98 (src/hbf.rs:258-292): the 98 dB tap set, same kernels with the other constants
#define HFS_NS hfs98
#define HFS_TAPS HbfTaps98
#define HFS_M hbf98_m
#include "hbf_fast_scalar_body.cuh"
#undef HFS_NS
#undef HFS_TAPS
#undef HFS_M

namespace idsp {


// Runs the tiled kernel over the first (n_out / TO) * TO frames of every lane.  Returns
// the number of frames it covered in *done (the caller finishes the tail with the
// generic kernel), or IDSP_HBF_FAST_NOT_APPLICABLE.
// (frame-major /32 with the LDGSTS gather measured slower than the generic thread-per-lane kernel, 823 vs 946
// GSa/s; with tensor-map input tiles the tiled kernel is the default for /32 too)
#define IDSP_DEF_DEC_FAST_TRY(NAME, NS) \
    static int NAME(idsp_ctx *ctx, int k, float *state, const float *x, float *y, size_t n_out, \
                                size_t lanes, size_t sstride, int layout, size_t *done) { \
        *done = 0; \
        const bool fm = layout == IDSP_FRAME_MAJOR; \
        if (ctx->policy == 1 || (fm && k > 4 && ctx->policy != 2 && getenv("IDSP_HBF_FM_LDGSTS"))) return IDSP_HBF_FAST_NOT_APPLICABLE; \
        const size_t TO = (size_t)NS::TT >> k; \
        const size_t ntiles = n_out / TO; \
        const bool ok = ntiles >= 1 && (((uintptr_t)x) & 15) == 0 && (fm || ((n_out << k) % 4) == 0); \
        if (!ok) return IDSP_HBF_FAST_NOT_APPLICABLE; \
        int r; \
        if (fm) { \
            switch (k) { \
                case 1: r = NS::launch<1, true>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
                case 2: r = NS::launch<2, true>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
                case 3: r = NS::launch<3, true>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
                case 4: r = NS::launch<4, true>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
                default: r = NS::launch<5, true>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
            } \
        } else { \
            switch (k) { \
                case 1: r = NS::launch<1, false>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
                case 2: r = NS::launch<2, false>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
                case 3: r = NS::launch<3, false>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
                case 4: r = NS::launch<4, false>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
                default: r = NS::launch<5, false>(ctx, state, x, y, n_out, ntiles, lanes, sstride); break; \
            } \
        } \
        if (r == IDSP_OK) *done = ntiles * TO; \
        return r; \
    }
IDSP_DEF_DEC_FAST_TRY(hbf_dec_fast_try_scalar, hfs)
IDSP_DEF_DEC_FAST_TRY(hbf98_dec_fast_try, hfs98)
#undef IDSP_DEF_DEC_FAST_TRY

}  // namespace idsp
