// tma_kernels.cuh -- TMA-pipelined lane kernels for 4-byte samples (i32 / f32).
//
// Same mapping as lane_kernels.cuh (one filter lane per thread, state in
// registers for the whole call), but the samples move through shared memory with
// the Tensor Memory Accelerator so the threads never wait on HBM:
//
//   * every warp owns 32 lanes and runs its own pipeline -- no CTA-wide barrier:
//     lane 0 issues `cp.async.bulk.tensor.2d` loads S-1 tiles ahead into a ring
//     of S shared-memory stages guarded by mbarriers (complete_tx), all 32 lanes
//     consume a stage, write results to one of O output stages and lane 0 hands
//     that stage to a TMA bulk store (bulk_group / wait_group.read).
//   * frame-major  flat[t*lanes + l]: box = [32 lanes x TF frames]; thread l reads
//     word l of every 128-byte row -> conflict-free LDS.32.
//   * lane-major   flat[l*frames + t]: box = [TF=16 frames x 32 lanes] written with
//     the 64-byte TMA swizzle; thread l reads its own 64-byte row as 4 x LDS.128
//     (chunk index XOR (l>>1)&3) -> conflict-free, 16 samples per 4 loads.
//   * out-of-range lanes / frames are zero-filled on load and clipped on store by
//     the TMA unit, so ragged shapes need no scalar tail code.
// Requirements: 16-byte aligned x/y, (lanes % 4 == 0) frame-major or
// (frames % 4 == 0) lane-major; otherwise the generic LDG kernels run.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "lane_kernels.cuh"

#define IDSP_TMA_NOT_APPLICABLE 12345

namespace idsp {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::
                     "l"(map),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <class T> struct Bits32;
template <> struct Bits32<int32_t> {
    __device__ __forceinline__ static int32_t from(uint32_t v) { return (int32_t)v; }
    __device__ __forceinline__ static uint32_t to(int32_t v) { return (uint32_t)v; }
};
template <> struct Bits32<float> {
    __device__ __forceinline__ static float from(uint32_t v) { return __uint_as_float(v); }
    __device__ __forceinline__ static uint32_t to(float v) { return __float_as_uint(v); }
};

// 4- or 8-byte sample <-> 32-bit words of a shared-memory tile (8-byte: low word first)
template <class T, int W = (int)sizeof(T) / 4> struct WordsOf;
template <class T> struct WordsOf<T, 1> {
    __device__ __forceinline__ static T from(const uint32_t *w) { return Bits32<T>::from(w[0]); }
    __device__ __forceinline__ static void to(T v, uint32_t *w) { w[0] = Bits32<T>::to(v); }
};
template <> struct WordsOf<int2, 2> {
    __device__ __forceinline__ static int2 from(const uint32_t *w) { return make_int2((int)w[0], (int)w[1]); }
    __device__ __forceinline__ static void to(int2 v, uint32_t *w) { w[0] = (uint32_t)v.x; w[1] = (uint32_t)v.y; }
};
template <> struct WordsOf<int64_t, 2> {
    __device__ __forceinline__ static int64_t from(const uint32_t *w) { return (int64_t)((uint64_t)w[0] | ((uint64_t)w[1] << 32)); }
    __device__ __forceinline__ static void to(int64_t v, uint32_t *w) { w[0] = (uint32_t)(uint64_t)v; w[1] = (uint32_t)((uint64_t)v >> 32); }
};
template <> struct WordsOf<double, 2> {
    __device__ __forceinline__ static double from(const uint32_t *w) { return __hiloint2double((int)w[1], (int)w[0]); }
    __device__ __forceinline__ static void to(double v, uint32_t *w) { w[0] = (uint32_t)__double2loint(v); w[1] = (uint32_t)__double2hiint(v); }
};

// ---------------------------------------------------------------- kernel
// LM = false: frame-major, tile [TF rows][32 words];  LM = true: lane-major,
// tile [32 rows (lanes)][16 words] with 64-byte swizzle (TF must be 16).
// Op::In and Op::Out are 4-byte types.  WPC warps per CTA, each independent.
// WIDE (frame-major only): the WPC warps of a CTA share one [32*WPC lanes x TF] box per
// tile (one TMA load / store per CTA tile, one __syncthreads per tile) so that a row
// of the box is 128*WPC contiguous bytes in HBM instead of 128.
template <class Op, bool LM, int TF, int S, int O, int WPC, bool WIDE = false>
__global__ void __launch_bounds__(WPC * 32)
tma_lanes_kernel(typename Op::Params p, const __grid_constant__ CUtensorMap mx,
                 const __grid_constant__ CUtensorMap my, size_t frames, size_t lanes,
                 size_t sstride) {
    using In = typename Op::In;
    using Out = typename Op::Out;
    static_assert((sizeof(In) == 4 || sizeof(In) == 8) && (sizeof(Out) == 4 || sizeof(Out) == 8), "4/8-byte in, 4/8-byte out");
    static_assert(!LM || TF == 16, "lane-major tiles are 16 frames (64B swizzle)");
    static_assert(!(WIDE && LM), "wide boxes are frame-major only");
    constexpr int IW = sizeof(In) / 4;              // input words per sample (Complex<i32>, i64, f64 = 2)
    constexpr int OW = sizeof(Out) / 4;             // output words per sample
    // lane-major 8-byte samples: the tile is [32 lanes][16 frames x 2 words] = 128-byte rows
    // written with the 128-byte TMA swizzle (chunk index XOR (l & 7))
    constexpr int BW = WIDE ? 32 * WPC : 32;        // box width in lanes
    constexpr int TILE_WORDS = TF * BW;             // samples per tile
    constexpr uint32_t TILE_IN_BYTES = TILE_WORDS * IW * 4;
    constexpr int NPIPE = WIDE ? 1 : WPC;            // independent pipelines per CTA
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int pipe = WIDE ? 0 : w;
    const int col = WIDE ? w * 32 + l : l;          // my column inside the box
    // per pipeline: S input stages, O output stages, S mbarriers
    uint32_t *wbase = reinterpret_cast<uint32_t *>(smem_raw) + (size_t)pipe * (S * IW + O * OW) * TILE_WORDS;
    uint32_t *sin = wbase;
    uint32_t *sout = wbase + S * IW * TILE_WORDS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NPIPE * (S * IW + O * OW) * TILE_WORDS * 4) + pipe * S;
    uint32_t *extra = reinterpret_cast<uint32_t *>(smem_raw + (size_t)NPIPE * (S * IW + O * OW) * TILE_WORDS * 4 + (size_t)NPIPE * S * 8);
    if constexpr (Op::SMEM_EXTRA_WORDS > 0) {
        Op::init_smem(p, extra, threadIdx.x, WPC * 32);
        __syncthreads();
    }

    const size_t box0 = WIDE ? (size_t)blockIdx.x * BW : ((size_t)blockIdx.x * WPC + w) * 32;
    if (!WIDE && box0 >= lanes) return;
    const size_t lane = box0 + col;
    const bool active = lane < lanes;
    const size_t ntiles = (frames + TF - 1) / TF;
    const bool leader = WIDE ? threadIdx.x == 0 : l == 0;
    auto sync_pipe = [&]() {
        if (WIDE) __syncthreads();
        else __syncwarp();
    };

    if (leader) {
#pragma unroll
        for (int s = 0; s < S; s++) mbar_init(smem_u32(&bars[s]), 1);
        mbar_fence_init();
    }
    sync_pipe();

    auto issue_load = [&](size_t tile) {
        const int s = (int)(tile % S);
        const uint32_t bar = smem_u32(&bars[s]);
        mbar_expect_tx(bar, TILE_IN_BYTES);
        if (LM)
            tma_load_2d(smem_u32(sin + s * TILE_WORDS * IW), &mx, (int)(tile * TF * IW), (int)box0, bar);
        else
            tma_load_2d(smem_u32(sin + s * TILE_WORDS * IW), &mx, (int)(box0 * IW), (int)(tile * TF), bar);
    };
    if (leader) {
#pragma unroll
        for (int s = 0; s < S - 1; s++)
            if ((size_t)s < ntiles) issue_load(s);
    }
    Op op;
    op.bind(p, extra);
    if (active) op.load(p, lane, sstride);

    for (size_t i = 0; i < ntiles; i++) {
        const int s = (int)(i % S);
        const int ob = (int)(i % O);
        if (leader) {
            // the output stage we are about to overwrite must have been read by its store
            tma_wait_read<O - 1>();
            if (i + S - 1 < ntiles) issue_load(i + S - 1);
        }
        sync_pipe();
        mbar_wait(smem_u32(&bars[s]), (uint32_t)((i / S) & 1));
        const uint32_t *tin = sin + s * TILE_WORDS * IW;
        uint32_t *tout = sout + ob * TILE_WORDS * OW;
        // frames beyond `frames` in the last tile are zero-filled by the TMA load and
        // clipped by the TMA store; they must not advance the filter state.
        const int nvalid = (int)((frames - i * TF) < (size_t)TF ? (frames - i * TF) : (size_t)TF);
        if constexpr (LM) {
            // row l = my lane: 4 * W chunks of 16 B (W words per sample); 64-byte rows: physical chunk =
            // c ^ ((l>>1)&3), 128-byte rows: c ^ (l & 7).  Four samples = IW input chunks -> OW output chunks.
            const uint4 *rin = reinterpret_cast<const uint4 *>(tin + l * 16 * IW);
            uint4 *rout = reinterpret_cast<uint4 *>(tout + l * 16 * OW);
            const int swi = IW == 2 ? (l & 7) : ((l >> 1) & 3);
            const int swo = OW == 2 ? (l & 7) : ((l >> 1) & 3);
            auto group = [&](int g, auto fast) {
                uint32_t wi[4 * IW], wo[4 * OW];
#pragma unroll
                for (int j = 0; j < IW; j++) {
                    const uint4 v = rin[(g * IW + j) ^ swi];
                    wi[4 * j] = v.x; wi[4 * j + 1] = v.y; wi[4 * j + 2] = v.z; wi[4 * j + 3] = v.w;
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    WordsOf<Out>::to(op_step<decltype(fast)::value>(op, p, WordsOf<In>::from(wi + q * IW)), wo + q * OW);
                }
#pragma unroll
                for (int j = 0; j < OW; j++)
                    rout[(g * OW + j) ^ swo] = make_uint4(wo[4 * j], wo[4 * j + 1], wo[4 * j + 2], wo[4 * j + 3]);
            };
            constexpr bool SPEC = op_speculative<Op>::value;  // see IDSP_LOCKIN_SPEC_MEMBERS (ops.cuh)
            if constexpr (SPEC) op.spec_begin();
            if (nvalid == TF) {
#pragma unroll
                for (int g = 0; g < 4; g++) group(g, std::integral_constant<bool, SPEC>());
            } else {  // frames % 4 == 0, so whole groups are valid or not
                for (int g = 0; g * 4 < nvalid; g++) group(g, std::integral_constant<bool, SPEC>());
            }
            if constexpr (SPEC) {
                if (op.spec_failed()) {  // per thread: redo this lane's tile with the exact step
                    op.spec_rollback();
#pragma unroll 1
                    for (int g = 0; g * 4 < nvalid; g++) group(g, std::false_type());
                }
            }
        } else {
            auto one = [&](int f, auto fast) {
                uint32_t wi[IW], wo[OW];
                if constexpr (IW == 2) {
                    const uint2 v = reinterpret_cast<const uint2 *>(tin)[f * BW + col];
                    wi[0] = v.x; wi[1] = v.y;
                } else {
                    wi[0] = tin[f * BW + col];
                }
                WordsOf<Out>::to(op_step<decltype(fast)::value>(op, p, WordsOf<In>::from(wi)), wo);
                if constexpr (OW == 2) reinterpret_cast<uint2 *>(tout)[f * BW + col] = make_uint2(wo[0], wo[1]);
                else tout[f * BW + col] = wo[0];
            };
            constexpr bool SPEC = op_speculative<Op>::value;  // see IDSP_LOCKIN_SPEC_MEMBERS (ops.cuh)
            if constexpr (SPEC) op.spec_begin();
            if (nvalid == TF) {
#pragma unroll
                for (int f = 0; f < TF; f++) one(f, std::integral_constant<bool, SPEC>());
            } else {
                for (int f = 0; f < nvalid; f++) one(f, std::integral_constant<bool, SPEC>());
            }
            if constexpr (SPEC) {
                if (op.spec_failed()) {  // per thread: redo this lane's tile with the exact step
                    op.spec_rollback();
#pragma unroll 1
                    for (int f = 0; f < nvalid; f++) one(f, std::false_type());
                }
            }
        }
        fence_async_smem();
        sync_pipe();
        if (leader) {
            if (LM)
                tma_store_2d(&my, smem_u32(tout), (int)(i * TF * OW), (int)box0);
            else
                tma_store_2d(&my, smem_u32(tout), (int)(box0 * OW), (int)(i * TF));
            tma_commit();
        }
    }
    if (leader) tma_wait_read<0>();
    if (active) op.store(p, lane, sstride);
}

// ---------------------------------------------------------------- lane-major, long rows
// Lane-major variant with longer contiguous runs per lane: a tile is NBOX boxes of
// [32 frames x 32 lanes] (128-byte rows, 128-byte TMA swizzle), i.e. 128*NBOX contiguous
// bytes per lane per tile instead of 64.  Thread l reads its own row of every box as
// 8 x LDS.128 with the chunk index XOR (l & 7) -> conflict free.  One warp per CTA.
template <class Op, int NBOX, int S, int O>
__global__ void __launch_bounds__(32)
tma_lanes_lm_kernel(typename Op::Params p, const __grid_constant__ CUtensorMap mx,
                    const __grid_constant__ CUtensorMap my, size_t frames, size_t lanes, size_t sstride) {
    using In = typename Op::In;
    using Out = typename Op::Out;
    static_assert(sizeof(In) == 4 && sizeof(Out) == 4, "4-byte samples only");
    constexpr int BOX_WORDS = 32 * 32;
    constexpr int TILE_WORDS = NBOX * BOX_WORDS;
    constexpr int TFT = 32 * NBOX;  // frames per tile
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int l = threadIdx.x;
    uint32_t *sin = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *sout = sin + S * TILE_WORDS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)(S + O) * TILE_WORDS * 4);
    const size_t lane0 = (size_t)blockIdx.x * 32;
    const size_t lane = lane0 + l;
    const bool active = lane < lanes;
    const size_t ntiles = (frames + TFT - 1) / TFT;
    if (l == 0) {
#pragma unroll
        for (int s = 0; s < S; s++) mbar_init(smem_u32(&bars[s]), 1);
        mbar_fence_init();
    }
    __syncwarp();
    auto issue_load = [&](size_t tile) {
        const int s = (int)(tile % S);
        const uint32_t bar = smem_u32(&bars[s]);
        mbar_expect_tx(bar, TILE_WORDS * 4);
#pragma unroll
        for (int b = 0; b < NBOX; b++)
            tma_load_2d(smem_u32(sin + s * TILE_WORDS + b * BOX_WORDS), &mx, (int)(tile * TFT + b * 32), (int)lane0, bar);
    };
    if (l == 0) {
#pragma unroll
        for (int s = 0; s < S - 1; s++)
            if ((size_t)s < ntiles) issue_load(s);
    }
    Op op;
    op.bind(p, nullptr);
    if (active) op.load(p, lane, sstride);
    const int sw = l & 7;
    for (size_t i = 0; i < ntiles; i++) {
        const int s = (int)(i % S);
        const int ob = (int)(i % O);
        if (l == 0) {
            tma_wait_read<O - 1>();
            if (i + S - 1 < ntiles) issue_load(i + S - 1);
        }
        __syncwarp();
        mbar_wait(smem_u32(&bars[s]), (uint32_t)((i / S) & 1));
        const size_t t0 = i * TFT;
        const int nvalid = (int)((frames - t0) < (size_t)TFT ? (frames - t0) : (size_t)TFT);
#pragma unroll
        for (int b = 0; b < NBOX; b++) {
            const uint4 *rin = reinterpret_cast<const uint4 *>(sin + s * TILE_WORDS + b * BOX_WORDS + l * 32);
            uint4 *rout = reinterpret_cast<uint4 *>(sout + ob * TILE_WORDS + b * BOX_WORDS + l * 32);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (b * 32 + c * 4 < nvalid) {  // frames % 4 == 0: whole chunks are valid or not
                    uint4 v = rin[c ^ sw];
                    uint4 r;
                    r.x = Bits32<Out>::to(op.step(p, Bits32<In>::from(v.x)));
                    r.y = Bits32<Out>::to(op.step(p, Bits32<In>::from(v.y)));
                    r.z = Bits32<Out>::to(op.step(p, Bits32<In>::from(v.z)));
                    r.w = Bits32<Out>::to(op.step(p, Bits32<In>::from(v.w)));
                    rout[c ^ sw] = r;
                }
            }
        }
        fence_async_smem();
        __syncwarp();
        if (l == 0) {
#pragma unroll
            for (int b = 0; b < NBOX; b++)
                tma_store_2d(&my, smem_u32(sout + ob * TILE_WORDS + b * BOX_WORDS), (int)(t0 + b * 32), (int)lane0);
            tma_commit();
        }
    }
    if (l == 0) tma_wait_read<0>();
    if (active) op.store(p, lane, sstride);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)f;
    }
    return fn;
}

// 2-D map over a row-major [rows][cols] array of 4-byte words
static bool make_map_2d(CUtensorMap *m, const void *base, uint64_t cols, uint64_t rows,
                        uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<void *>(base), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <class Op, bool LM, int TF, int S, int O, int WPC, bool WIDE = false>
static int tma_launch_cfg(idsp_ctx *ctx, const typename Op::Params &p, const void *x, void *y,
                          size_t frames, size_t lanes, size_t sstride) {
    CUtensorMap mx, my;
    bool ok;
    constexpr int BW = WIDE ? 32 * WPC : 32;
    constexpr int IW = sizeof(typename Op::In) / 4;
    constexpr int OW = sizeof(typename Op::Out) / 4;
    static_assert(BW * OW <= 256 && BW * IW <= 256, "TMA box dimension limit");
    if (LM) {
        ok = make_map_2d(&mx, x, frames * IW, lanes, TF * IW, 32,
                         IW == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B) &&
             make_map_2d(&my, y, frames * OW, lanes, TF * OW, 32,
                         OW == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    } else {
        ok = make_map_2d(&mx, x, lanes * IW, frames, BW * IW, TF, CU_TENSOR_MAP_SWIZZLE_NONE) &&
             make_map_2d(&my, y, lanes * OW, frames, BW * OW, TF, CU_TENSOR_MAP_SWIZZLE_NONE);
    }
    if (!ok) return IDSP_TMA_NOT_APPLICABLE;
    constexpr int NPIPE = WIDE ? 1 : WPC;
    constexpr size_t smem = (size_t)NPIPE * (S * IW + O * OW) * TF * BW * 4 + (size_t)NPIPE * S * 8 +
                            (size_t)Op::SMEM_EXTRA_WORDS * 4;
    auto kern = tma_lanes_kernel<Op, LM, TF, S, O, WPC, WIDE>;
    IDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t warps = (lanes + 31) / 32;
    unsigned grid = (unsigned)((warps + WPC - 1) / WPC);
    kern<<<grid, WPC * 32, smem, ctx->stream>>>(p, mx, my, frames, lanes, sstride);
    IDSP_KERNEL_FAMILY(ctx, LM ? "tma lane-major" : (WIDE ? "tma frame-major wide" : "tma frame-major"));
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

// Resident CTAs per SM of one frame-major configuration (cached) and the cost of its last,
// partially filled wave: a CTA owns its lanes for the whole time axis, so with W waves of
// CTAs the launch lasts ~ceil(W) rounds while doing W rounds of work.  Compute-bound ops pick
// the configuration with the smallest ceil(W) / W (HBM-bound ops do not care).
template <class Op, bool LM, int TF, int S, int O, int WPC, bool WIDE>
static int tma_cfg_occupancy() {
    static int occ = -1;
    if (occ < 0) {
        constexpr int BW = WIDE ? 32 * WPC : 32;
        constexpr int IW = sizeof(typename Op::In) / 4;
        constexpr int OW = sizeof(typename Op::Out) / 4;
        constexpr int NPIPE = WIDE ? 1 : WPC;
        constexpr size_t smem = (size_t)NPIPE * (S * IW + O * OW) * TF * BW * 4 + (size_t)NPIPE * S * 8 +
                                (size_t)Op::SMEM_EXTRA_WORDS * 4;
        auto kern = tma_lanes_kernel<Op, LM, TF, S, O, WPC, WIDE>;
        int n = 0;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, WPC * 32, smem) != cudaSuccess || n < 1)
            n = 1;
        occ = n;
    }
    return occ;
}
static inline double tail_cost(size_t ctas, size_t sms, int occ) {
    const double w = (double)ctas / (double)(sms * (size_t)occ);
    const double full = (double)(size_t)w;
    return (w > full ? full + 1.0 : full) / w;
}

template <class Op, int NBOX, int S, int O>
static int tma_launch_lm(idsp_ctx *ctx, const typename Op::Params &p, const void *x, void *y,
                         size_t frames, size_t lanes, size_t sstride) {
    CUtensorMap mx, my;
    if (!(make_map_2d(&mx, x, frames, lanes, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B) &&
          make_map_2d(&my, y, frames, lanes, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)))
        return IDSP_TMA_NOT_APPLICABLE;
    constexpr size_t smem = (size_t)(S + O) * NBOX * 4096 + S * 8;
    auto kern = tma_lanes_lm_kernel<Op, NBOX, S, O>;
    IDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)((lanes + 31) / 32);
    kern<<<grid, 32, smem, ctx->stream>>>(p, mx, my, frames, lanes, sstride);
    IDSP_KERNEL_FAMILY(ctx, "tma lane-major long rows");
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

// Lane-major default: 2 boxes of 32 frames per tile (256 contiguous bytes per lane), 3 load
// and 2 store stages (tools/sweep_biquad.py SWEEP_LM=1: 708 GSa/s vs 554 for 64-byte rows);
// short streams use the 16-frame kernel.
template <class Op>
static int tma_launch_lm_auto(idsp_ctx *ctx, const typename Op::Params &p, const void *x, void *y,
                              size_t frames, size_t lanes, size_t sstride) {
    if constexpr (Op::HEAVY) {
        // compute-bound ops: 12 KB (128-byte rows) or 8 KB (64-byte rows) per warp instead of 40 KB, so
        // 18..28 warps are resident per SM instead of 5 (Cascade<4> i32 128 -> 316 GSa/s, PLL 324 -> 515)
        if (frames >= 128) return tma_launch_lm<Op, 1, 2, 1>(ctx, p, x, y, frames, lanes, sstride);
        return tma_launch_cfg<Op, true, 16, 3, 1, 1>(ctx, p, x, y, frames, lanes, sstride);
    }
    // ops with 64-bit state arithmetic between HBM- and compute-bound: 2 load + 1 store stage of the same
    // 256-byte rows (24 KB, 9 warps per SM instead of 5): Lowpass<2> 557 -> 656 GSa/s, Lowpass<1> 657 -> 679
    if constexpr (Op::LM_SMALL) {
        if (frames >= 128) return tma_launch_lm<Op, 2, 2, 1>(ctx, p, x, y, frames, lanes, sstride);
    }
    if (frames >= 128) return tma_launch_lm<Op, 2, 3, 2>(ctx, p, x, y, frames, lanes, sstride);
    return tma_launch_cfg<Op, true, 16, 4, 2, 1>(ctx, p, x, y, frames, lanes, sstride);
}

// Frame-major default: tiles of 8 frames, 4 load stages, 2 store stages, and the widest
// box (32*WPC lanes, WPC <= 8) that still leaves one CTA per SM.  Measured on the
// 65 536-lane i32 DF1 stream (tools/sweep_biquad.py, profiles/sweep_biquad_r1.md):
// 128-byte rows (WPC 1) 4.6 TB/s, 1 KB rows (WPC 8) 6.5 TB/s = the copy peak.
template <class Op>
static int tma_launch_fm_auto(idsp_ctx *ctx, const typename Op::Params &p, const void *x, void *y,
                              size_t frames, size_t lanes, size_t sstride) {
    const size_t sms = (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 148);
    auto ctas = [&](size_t wpc) { return (lanes + 32 * wpc - 1) / (32 * wpc); };
    // compute-bound ops: independent per-warp pipelines of 16-frame tiles (no CTA-wide barrier per tile);
    // Cascade<4> i32 276 -> 302 GSa/s, PLL 455 -> 487 against the wide boxes that serve the HBM-bound ops
    if constexpr (Op::HEAVY) return tma_launch_cfg<Op, false, 16, 4, 2, 1>(ctx, p, x, y, frames, lanes, sstride);
    if (ctas(8) >= sms) return tma_launch_cfg<Op, false, 8, 4, 2, 8, true>(ctx, p, x, y, frames, lanes, sstride);
    if (ctas(4) >= sms) return tma_launch_cfg<Op, false, 8, 4, 2, 4, true>(ctx, p, x, y, frames, lanes, sstride);
    if (ctas(2) >= sms) return tma_launch_cfg<Op, false, 8, 4, 2, 2, true>(ctx, p, x, y, frames, lanes, sstride);
    return tma_launch_cfg<Op, false, 16, 4, 2, 1>(ctx, p, x, y, frames, lanes, sstride);
}

// Returns IDSP_TMA_NOT_APPLICABLE when the generic kernels must be used.
template <class Op>
static int tma_try_launch(idsp_ctx *ctx, const typename Op::Params &p, const typename Op::In *x,
                          typename Op::Out *y, size_t frames, size_t lanes, size_t sstride,
                          int layout) {
    static_assert((sizeof(typename Op::In) == 4 || sizeof(typename Op::In) == 8) &&
                  (sizeof(typename Op::Out) == 4 || sizeof(typename Op::Out) == 8), "");
    constexpr bool OUT8 = sizeof(typename Op::Out) == 8 || sizeof(typename Op::In) == 8;  // any 8-byte side
    if (ctx->policy == 1) return IDSP_TMA_NOT_APPLICABLE;
    const bool lm = layout == IDSP_LANE_MAJOR;
    bool ok = (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && frames >= 16 &&
              frames < (1ull << 30) && lanes < (1ull << 30) &&
              (lm ? (frames % 4 == 0) : (lanes % 4 == 0));
    if (!ok) {
        if (ctx->policy == 2) {
            idsp_set_error("TMA kernels forced but shape/alignment does not qualify");
            return IDSP_EINVAL;
        }
        return IDSP_TMA_NOT_APPLICABLE;
    }
    int r;
    if constexpr (OUT8) {
        // 4-byte in / 8-byte out (lock-in): 128-lane boxes (the 8-byte box row is 256 words)
        const size_t sms = (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 148);
        if (lm) {
            // one warp per CTA, 64-byte input rows / 128-byte output rows, 10 KB per warp: the op
            // is ALU-bound, so resident warps count for more than long DRAM bursts
            if constexpr (Op::SMEM_EXTRA_WORDS >= 1024)  // a big per-CTA table is shared between 4 warps
                r = tma_launch_cfg<Op, true, 16, 3, 1, 4>(ctx, p, x, y, frames, lanes, sstride);
            else
                r = tma_launch_cfg<Op, true, 16, 3, 1, 1>(ctx, p, x, y, frames, lanes, sstride);
#ifdef IDSP_TUNE
        } else if (getenv("IDSP_OUT8_CFG") && (Op::HEAVY || sizeof(typename Op::In) == 8)) {
            // tuning builds: tile shape / residency sweep of the compute-bound 8-byte ops (lock-in)
            switch (atoi(getenv("IDSP_OUT8_CFG"))) {
                case 1: r = tma_launch_cfg<Op, false, 4, 4, 2, 4, true>(ctx, p, x, y, frames, lanes, sstride); break;
                case 2: r = tma_launch_cfg<Op, false, 4, 3, 2, 4, true>(ctx, p, x, y, frames, lanes, sstride); break;
                case 3: r = tma_launch_cfg<Op, false, 8, 3, 1, 1>(ctx, p, x, y, frames, lanes, sstride); break;
                case 4: r = tma_launch_cfg<Op, false, 8, 3, 1, 2>(ctx, p, x, y, frames, lanes, sstride); break;
                case 5: r = tma_launch_cfg<Op, false, 4, 4, 2, 2>(ctx, p, x, y, frames, lanes, sstride); break;
                case 6: r = tma_launch_cfg<Op, false, 8, 2, 1, 4, true>(ctx, p, x, y, frames, lanes, sstride); break;
                case 7: r = tma_launch_cfg<Op, false, 4, 3, 1, 4, true>(ctx, p, x, y, frames, lanes, sstride); break;
                case 8: r = tma_launch_cfg<Op, false, 16, 2, 1, 1>(ctx, p, x, y, frames, lanes, sstride); break;
                case 9: r = tma_launch_cfg<Op, false, 4, 4, 2, 1>(ctx, p, x, y, frames, lanes, sstride); break;
                case 10: r = tma_launch_cfg<Op, false, 8, 3, 2, 2, true>(ctx, p, x, y, frames, lanes, sstride); break;
                case 11: r = tma_launch_cfg<Op, false, 4, 4, 2, 2, true>(ctx, p, x, y, frames, lanes, sstride); break;
                default: r = tma_launch_cfg<Op, false, 8, 3, 2, 4, true>(ctx, p, x, y, frames, lanes, sstride); break;
            }
#endif
        } else if (Op::HEAVY || std::is_integral<typename Op::In>::value) {
            // compute-bound ops with an 8-byte side (lock-in: ALU pipe; FM discriminator: atan2; i64 biquad:
            // 64 x 64 -> 128 multiplies): independent per-warp pipelines of 16-frame tiles, 2 load stages +
            // 1 store stage, no CTA-wide barrier per tile.  Sweeps of 12 tile shapes / residencies on 131 072
            // and 65 536 lanes (profiles/r2_sweep_lockin_tile_shapes.log, r2_sweep_8byte_tile_shapes.log):
            // lock-in 336 GSa/s against 326 for the 128-lane shared boxes, (x, phase) lock-in 269 / 254,
            // FM discriminator 188 / 180, i64 DF1 145 / 137; more resident warps do not help (the kernels
            // are bound by a pipe, not by latency).
            // (an Op with a big per-CTA table -- the replicated cossin table -- shares it between 4 warps)
            if constexpr (Op::SMEM_EXTRA_WORDS >= 1024)
                r = tma_launch_cfg<Op, false, 16, 2, 1, 4>(ctx, p, x, y, frames, lanes, sstride);
            else
                r = tma_launch_cfg<Op, false, 16, 2, 1, 1>(ctx, p, x, y, frames, lanes, sstride);
        } else if ((lanes + 127) / 128 >= sms) {
            // HBM-bound 8-byte streams (f64 biquads): 128-lane boxes (1 KB rows) with 2 load stages + 1 store
            // stage -- 24 KB per CTA instead of 48, i.e. twice the CTAs in flight: f64 DF1 347 -> 389 GSa/s
            // = 95 % of the HBM peak (same sweep)
            r = tma_launch_cfg<Op, false, 8, 2, 1, 4, true>(ctx, p, x, y, frames, lanes, sstride);
        } else
            r = tma_launch_cfg<Op, false, 8, 4, 2, 1>(ctx, p, x, y, frames, lanes, sstride);
    } else if (lm) {
#ifdef IDSP_TUNE
        const char *e = getenv("IDSP_TMA_LMCFG");
        int cfg = (e && Op::TUNABLE) ? atoi(e) : -1;
        switch (cfg) {
            case 0: r = tma_launch_cfg<Op, true, 16, 4, 2, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 1: r = tma_launch_lm<Op, 1, 3, 2>(ctx, p, x, y, frames, lanes, sstride); break;
            case 2: r = tma_launch_lm<Op, 2, 3, 2>(ctx, p, x, y, frames, lanes, sstride); break;
            case 3: r = tma_launch_lm<Op, 4, 2, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 4: r = tma_launch_lm<Op, 4, 2, 2>(ctx, p, x, y, frames, lanes, sstride); break;
            case 5: r = tma_launch_lm<Op, 2, 2, 2>(ctx, p, x, y, frames, lanes, sstride); break;
            case 6: r = tma_launch_lm<Op, 1, 4, 2>(ctx, p, x, y, frames, lanes, sstride); break;
            case 7: r = tma_launch_lm<Op, 8, 2, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 8: r = tma_launch_lm<Op, 2, 4, 2>(ctx, p, x, y, frames, lanes, sstride); break;
            case 9: r = tma_launch_lm<Op, 2, 3, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 10: r = tma_launch_lm<Op, 3, 3, 2>(ctx, p, x, y, frames, lanes, sstride); break;
            case 11: r = tma_launch_lm<Op, 3, 2, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 12: r = tma_launch_lm<Op, 3, 3, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            default: r = tma_launch_lm_auto<Op>(ctx, p, x, y, frames, lanes, sstride); break;
        }
#else
        r = tma_launch_lm_auto<Op>(ctx, p, x, y, frames, lanes, sstride);
#endif
    } else {
#ifdef IDSP_TUNE
        // tuning builds only (tools/sweep_biquad.py): pick a tile configuration at run time
        const char *e = getenv("IDSP_TMA_CFG");
        int cfg = (e && Op::TUNABLE) ? atoi(e) : -1;
        switch (cfg) {
            case 1: r = tma_launch_cfg<Op, false, 16, 6, 2, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 2: r = tma_launch_cfg<Op, false, 32, 3, 2, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 3: r = tma_launch_cfg<Op, false, 32, 4, 2, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 4: r = tma_launch_cfg<Op, false, 8, 8, 3, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 5: r = tma_launch_cfg<Op, false, 16, 4, 2, 2>(ctx, p, x, y, frames, lanes, sstride); break;
            case 6: r = tma_launch_cfg<Op, false, 16, 4, 2, 4>(ctx, p, x, y, frames, lanes, sstride); break;
            case 7: r = tma_launch_cfg<Op, false, 16, 4, 2, 2, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 8: r = tma_launch_cfg<Op, false, 16, 4, 2, 4, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 9: r = tma_launch_cfg<Op, false, 16, 4, 2, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 10: r = tma_launch_cfg<Op, false, 8, 6, 2, 4, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 11: r = tma_launch_cfg<Op, false, 32, 3, 2, 2, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 12: r = tma_launch_cfg<Op, false, 16, 3, 1, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            case 13: r = tma_launch_cfg<Op, false, 16, 6, 3, 4, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 14: r = tma_launch_cfg<Op, false, 8, 8, 4, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 15: r = tma_launch_cfg<Op, false, 16, 3, 2, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 16: r = tma_launch_cfg<Op, false, 16, 5, 2, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 17: r = tma_launch_cfg<Op, false, 32, 3, 2, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 18: r = tma_launch_cfg<Op, false, 8, 6, 2, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 19: r = tma_launch_cfg<Op, false, 16, 4, 3, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 20: r = tma_launch_cfg<Op, false, 8, 4, 2, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 21: r = tma_launch_cfg<Op, false, 4, 8, 4, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 22: r = tma_launch_cfg<Op, false, 16, 2, 2, 8, true>(ctx, p, x, y, frames, lanes, sstride); break;
            case 0: r = tma_launch_cfg<Op, false, 16, 4, 2, 1>(ctx, p, x, y, frames, lanes, sstride); break;
            default: r = tma_launch_fm_auto<Op>(ctx, p, x, y, frames, lanes, sstride); break;
        }
#else
        r = tma_launch_fm_auto<Op>(ctx, p, x, y, frames, lanes, sstride);
#endif
    }
    if (r == IDSP_TMA_NOT_APPLICABLE && ctx->policy == 2) {
        idsp_set_error("TMA kernels forced but cuTensorMapEncodeTiled is unavailable");
        return IDSP_EINVAL;
    }
    return r;
}

// TMA kernel when the Op has 4-byte samples and the shape qualifies, generic kernel otherwise
template <class Op>
static int launch_lanes_best(idsp_ctx *ctx, const typename Op::Params &p, const typename Op::In *x,
                             typename Op::Out *y, size_t frames, size_t lanes, size_t sstride, int layout) {
    if constexpr ((sizeof(typename Op::In) == 4 || sizeof(typename Op::In) == 8) &&
                  (sizeof(typename Op::Out) == 4 || sizeof(typename Op::Out) == 8)) {
        int tr = tma_try_launch<Op>(ctx, p, x, y, frames, lanes, sstride, layout);
        if (tr != IDSP_TMA_NOT_APPLICABLE) return tr;
    }
    return launch_lanes<Op>(ctx, p, x, y, frames, lanes, sstride, layout);
}

}  // namespace idsp
