// hbf_int_fast_body.cuh -- body of the tiled interpolator, included by hbf_int_fast.cuh once per tile shape
// (HFI_NS = namespace, HFI_NL = lanes per CTA, HFI_TOUT = output samples per lane and tile) and tap set
// (HFI_TAPS = tap struct template, HFI_M = tap count function).  No include guard.
namespace idsp {
namespace HFI_NS {

constexpr int NL = HFI_NL;      // lanes per CTA
#ifdef HFI_NT
constexpr int NT = HFI_NT;      // threads per CTA (FIR warps)
constexpr int MINB = HFI_MINB;  // CTAs per SM the kernel is compiled for
#else
constexpr int NT = 128;
constexpr int MINB = 4;
#endif
constexpr int TOUT = HFI_TOUT;  // output samples per lane per tile

__host__ __device__ constexpr int up4(int v) { return (v + 3) & ~3; }
__host__ __device__ constexpr int oddpitch(int v) { return (up4(v) / 4) % 2 ? up4(v) : up4(v) + 4; }
// stage s of a x2^K cascade uses TAPS[s] (lowest rate first, src/hbf.rs:503-512)
__host__ __device__ constexpr int st_m(int s) { return HFI_M(s); }
__host__ __device__ constexpr int ti(int K) { return TOUT >> K; }                 // inputs per tile
__host__ __device__ constexpr int st_nin(int K, int s) { return ti(K) << s; }      // inputs of stage s per tile
__host__ __device__ constexpr int st_r(int K, int s) {                             // inputs per item
    return st_nin(K, s) / 16 >= 8 ? 8 : 4;
}
__host__ __device__ constexpr int hist(int s) { return up4(2 * st_m(s) - 1); }
__host__ __device__ constexpr int pitch(int K, int s) { return oddpitch(hist(s) + st_nin(K, s)); }
__host__ __device__ constexpr int off_u(int K, int s) {  // float offset of stage s's input rows
    int o = 0;
    for (int i = 0; i < s; i++) o += NL * pitch(K, i);
    return o;
}
constexpr int OUT_PITCH = oddpitch(TOUT);
__host__ __device__ constexpr int off_out(int K) { return off_u(K, K); }
// Frame-major x16 output through the tensor-map unit (mirror image of FmTma in hbf_fast_scalar_body.cuh): the
// staged tile of a lane pair is 32 lines of 128 bytes ([frame][pair][16 floats], 128-byte swizzle by absolute
// shared-memory address) plus one unused line, so that the four pairs' swizzle phases stagger and the eight lanes
// of a quarter-warp store conflict free; four `cp.async.bulk.tensor` stores move a CTA's tile.
template <int K> struct FmOut {
    static constexpr int R = 1 << K;       // floats per output frame and lane
    static constexpr int LPL = 32 / R;     // lanes per 128-byte line
    static constexpr int NB = NL / LPL;    // boxes per staging buffer
    static constexpr int FT = TOUT / R;    // frames per tile
    static constexpr int LINES = FT + 1 - FT % 2;  // odd box pitch: the boxes' swizzle phases stagger
    static constexpr int BUF_LINES = NB * LINES;
    // address of chunk 0 of (buffer, lane, frame); chunk c of that frame sits at the result ^ (c << 4)
    __device__ __forceinline__ static uint32_t frame_base(uint32_t stg, int buf, int lane, int frame) {
        const uint32_t line = stg + (uint32_t)(buf * BUF_LINES + (lane / LPL) * LINES + frame) * 128u;
        return line + ((((uint32_t)(lane % LPL) * (R / 4)) ^ (line >> 7)) & 7u) * 16u;
    }
};
__host__ __device__ constexpr bool fm_out_rate(int K) { return K >= 2 && K <= 5 && NL == 8 && TOUT == 512; }
__host__ __device__ constexpr int stage_floats(int K) {
    const int R = 1 << K, lpl = 32 / (R > 32 ? 32 : R), ft = TOUT / R;
    const int rows = 2 * NL * OUT_PITCH, boxes = 2 * (NL / (lpl ? lpl : 1)) * (ft + 1 - ft % 2) * 32;
    return (fm_out_rate(K) && boxes > rows) ? boxes : rows;
}
__host__ __device__ constexpr int smem_floats(int K) { return off_out(K) + stage_floats(K); }
__host__ __device__ constexpr size_t smem_bytes(int K) { return (size_t)smem_floats(K) * 4; }
__host__ __device__ constexpr int st_word(int s) {  // ABI state word offset of stage s
    int w = 0;
    for (int i = 0; i < s; i++) w += 2 * st_m(i) - 1;
    return w;
}


// One item: inputs n0 .. n0+R-1 of row `row` = [H hist | n new]; writes 2R outputs to dst
template <int TI_, int R, bool FMT = false> struct IntItem {
    static constexpr int M = HFI_TAPS<TI_>::M;
    static constexpr int LEN = 2 * M - 1;
    static constexpr int H = up4(LEN);
    static constexpr int RO = H - LEN;
    static constexpr int W = up4(RO + R + LEN);
    __device__ __forceinline__ static void run(const float *row, int n0, float *dst, const uint32_t *dst_sw = nullptr) {
        float w[W];
#pragma unroll
        for (int j = 0; j < W / 4; j++) {
            float4 v = lds128v(row + n0 + 4 * j);
            w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
        }
        float o[2 * R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            // window of input n0+q: w[RO+q .. RO+q+2M-1]
            float acc = (w[RO + q + 2 * M - 1] + w[RO + q]) * HFI_TAPS<TI_>::c(0);
#pragma unroll
            for (int i = 1; i < M; i++)
                acc = acc + (w[RO + q + 2 * M - 1 - i] + w[RO + q + i]) * HFI_TAPS<TI_>::c(i);
            o[2 * q] = acc;
            o[2 * q + 1] = w[RO + q + M];
        }
#pragma unroll
        for (int j = 0; j < 2 * R / 4; j++) {
            if constexpr (FMT) {  // dst_sw[j] = swizzled shared-memory address of the item's chunk j (FmOut)
                static_assert(!FMT || R == 8, "an item is 16 outputs = four chunks");
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst_sw[j]), "f"(o[4 * j]),
                             "f"(o[4 * j + 1]), "f"(o[4 * j + 2]), "f"(o[4 * j + 3])
                             : "memory");
            } else {
                reinterpret_cast<float4 *>(dst)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
        }
    }
};

// Move the tails of the [hist | n new] input rows of stage s (all NL lanes) to their heads
// (copy_within, src/hbf.rs:231).  One 16-byte piece per thread, a whole row inside one warp:
// everything is read before anything is written (head and tail overlap when hist > n).
template <int K, int s>
__device__ __forceinline__ void carry_rows(float *sm, int warp, int lid) {
    constexpr int C = hist(s) / 4;
    static_assert(C <= 32, "row history too long for one warp");
    // a lane's C pieces sit in a group of CP = 2^k >= C consecutive threads (see hbf_fast_scalar_body.cuh)
    constexpr int CP = C <= 1 ? 1 : C <= 2 ? 2 : C <= 4 ? 4 : C <= 8 ? 8 : C <= 16 ? 16 : 32;
    constexpr int LP = 32 / CP;  // lanes per warp pass
    const int sub = lid / CP, j = lid % CP;
    for (int l0 = warp * LP; l0 < NL; l0 += (NT / 32) * LP) {  // l0 is warp-uniform
        const int lane = l0 + sub;
        const bool act = j < C && lane < NL;
        float *row = sm + off_u(K, s) + lane * pitch(K, s) + 4 * j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) v = lds128v(row + st_nin(K, s));
        __syncwarp();
        if (act) *reinterpret_cast<float4 *>(row) = v;
        __syncwarp();
    }
}

template <int K, int s, bool FMT = false> struct StageRun {
    __device__ __forceinline__ static void run(float *sm, int tid, int obuf) {
        constexpr int R = st_r(K, s);
        constexpr int ITEMS = NL * st_nin(K, s) / R;
        const float *U = sm + off_u(K, s);
        for (int idx = tid; idx < ITEMS; idx += NT) {
            const int lane = idx % NL, n0 = (idx / NL) * R;
            float *dst;
            if constexpr (s == K - 1) dst = sm + off_out(K) + (obuf * NL + lane) * OUT_PITCH + 2 * n0;
            else dst = sm + off_u(K, s + 1) + lane * pitch(K, s + 1) + hist(s + 1) + 2 * n0;
            if constexpr (FMT && s == K - 1) {  // chunk j = outputs 2 * n0 + 4j ..: frame (..) / 2^K, chunk ((..) % 2^K) / 4
                uint32_t d[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int o0 = 2 * n0 + 4 * j;
                    d[j] = FmOut<K>::frame_base(smem_u32(sm + off_out(K)), obuf, lane, o0 >> K) ^
                           (uint32_t)(((o0 & ((1 << K) - 1)) / 4) << 4);
                }
                IntItem<s, R, true>::run(U + lane * pitch(K, s), n0, nullptr, d);
            }
            else
                IntItem<s, R>::run(U + lane * pitch(K, s), n0, dst);
        }
        // the input rows of the previous stage were consumed one barrier ago
        if constexpr (s >= 1) carry_rows<K, s - 1>(sm, tid >> 5, tid & 31);
    }
};

template <int K, int s, bool LOAD> struct StateIO {
    __device__ __forceinline__ static void run(float *sm, float *st, size_t sstride, size_t lane0, int nl, int tid) {
        if constexpr (s < K) {
            constexpr int LEN = 2 * st_m(s) - 1;
            float *stw = st + (size_t)st_word(s) * sstride + lane0;
            if constexpr (LOAD) {  // loads first, then stores: one global round trip (see hbf_fast_scalar_body.cuh)
                constexpr int IT = (LEN * NL + NT - 1) / NT;
                float v[IT];
#pragma unroll
                for (int it = 0; it < IT; it++) {
                    const int idx = tid + it * NT, lane = idx % NL, w = idx / NL;
                    v[it] = (idx < LEN * NL && lane < nl) ? __ldg(stw + (size_t)w * sstride + lane) : 0.f;
                }
#pragma unroll
                for (int it = 0; it < IT; it++) {
                    const int idx = tid + it * NT, lane = idx % NL, w = idx / NL;
                    if (idx < LEN * NL && lane < nl) sm[off_u(K, s) + lane * pitch(K, s) + (hist(s) - LEN + w)] = v[it];
                }
            } else {
                for (int idx = tid; idx < LEN * NL; idx += NT) {
                    const int lane = idx % NL, w = idx / NL;
                    if (lane >= nl) continue;
                    stw[(size_t)w * sstride + lane] = sm[off_u(K, s) + lane * pitch(K, s) + (hist(s) - LEN + w)];
                }
            }
            StateIO<K, s + 1, LOAD>::run(sm, st, sstride, lane0, nl, tid);
        }
    }
};

// FM = false: x, y lane-major.  FM = true: frame-major x[t][lane], y[t][lane][2^K]: the
// input tile is prefetched element-wise and the staged output rows leave as 16-byte (x2: 8-byte) pieces
// (8 lanes x 4 pieces of one frame per warp store = up to 512 contiguous bytes of HBM).
// BQ = true (lane-major only): a fifth warp runs an iir::Biquad DF1 f32 (src/iir/biquad.rs:366-383) over every
// staged output tile in place before it is stored -- one thread per lane, the recurrence is serial in
// time -- while the four FIR warps already compute the next tile (HbfInt -> Biquad of the config-5 chain
// without a second pass over HBM).
template <int K, bool FM, bool BQ = false, bool FMT = false>
__global__ void __launch_bounds__(NT + (BQ ? 32 : 0), BQ ? (NT > 128 ? MINB : HFI_BQ_MINB) : MINB)
hbf_int_fast_kernel(float *st, const float *x, float *y, size_t n_in, size_t ntiles, size_t lanes, size_t sstride,
                    Df1Op<float, false>::Params bq, const __grid_constant__ CUtensorMap ymap) {
    static_assert(!FMT || (FM && !BQ && fm_out_rate(K) && 2 * FmOut<FMT ? K : 4>::BUF_LINES * 32 <= stage_floats(K)),
                  "tensor-map output: frame-major x4 ... x32, 8 lanes x 512 samples");
    static_assert(!(BQ && FM), "the fused biquad variant is lane-major");
    constexpr int NTA = NT + (BQ ? 32 : 0);
    auto fir_sync = [&]() {
        if constexpr (BQ) nbar_sync(1, NT);
        else __syncthreads();
    };
    constexpr int TI = ti(K);
    constexpr int NV = FM ? NL * TI : NL * TI / 4;  // loads per input tile (floats if FM, float4 else)
    constexpr int NVT = (NV + NT - 1) / NT;         // ... per thread
    extern __shared__ __align__(1024) float sm[];
    const int tid = threadIdx.x;
    const size_t lane0 = (size_t)blockIdx.x * NL;
    const int nl = (int)((lanes - lane0) < (size_t)NL ? (lanes - lane0) : (size_t)NL);
    const size_t n_out = n_in << K;  // row stride of y in floats

    for (int i = tid; i < smem_floats(K); i += NTA) sm[i] = 0.f;
    __syncthreads();
    if constexpr (BQ) {
        if (tid >= NT) {
            // ---- the biquad warp: thread j owns lane j of the CTA (NL <= 32).  The recurrence is a chain
            // of dependent FP32 operations, so its cost per tile does not depend on how many of the 32
            // threads are in use: more lanes per CTA (hfi16) make it cheaper per sample.
            static_assert(NL <= 32, "one biquad warp");
            const int j = tid - NT;
            const bool act = j < nl && j < NL;
            Df1Op<float, false> op;
            op.x1 = op.x2 = op.y1 = op.y2 = 0.f;
            if (act) op.load(bq, lane0 + j, sstride);
            // The bulk store of tile i is not waited for where it is issued: its shared-memory reads complete
            // while the first quarter of tile i + 1 is filtered, and only then is buffer i & 1 handed back to
            // the FIR warps (they need it for tile i + 2, one and a half tiles later) -- the store latency
            // (11 % of this warp's time, and this warp is the critical path of the CTA) leaves the chain.
            constexpr int Q1 = TOUT / 16;  // pieces of 4 samples filtered before the previous buffer is released
            for (size_t i = 0; i < ntiles; i++) {
                const int ob = (int)(i & 1);
                nbar_sync(2 + ob, NTA);  // tile i is staged
                float *row = sm + off_out(K) + (ob * NL + (act ? j : 0)) * OUT_PITCH;
                float4 v0 = lds128v(row), v1 = lds128v(row + 4);  // loads run two pieces ahead of the chain
                auto piece = [&](int q) {
                    float4 v = v0;
                    v0 = v1;
                    if (q + 2 < TOUT / 4) v1 = lds128v(row + 4 * (q + 2));
                    v.x = op.step(bq, v.x);
                    v.y = op.step(bq, v.y);
                    v.z = op.step(bq, v.z);
                    v.w = op.step(bq, v.w);
                    if (act) *reinterpret_cast<float4 *>(row + 4 * q) = v;
                };
#pragma unroll 4
                for (int q = 0; q < Q1; q++) piece(q);
                if (i >= 1 && i + 1 < ntiles) {  // tile i - 1 has left its buffer: the FIR warps may stage tile i + 1
                    if (act) tma_wait_read<0>();
                    __syncwarp();
                    nbar_arrive(4 + (ob ^ 1), NTA);
                }
#pragma unroll 4
                for (int q = Q1; q < TOUT / 4; q++) piece(q);
                if (act) {
                    fence_async_smem();
                    bulk_store_1d(y + (lane0 + j) * n_out + i * TOUT, smem_u32(row), TOUT * 4);
                    tma_commit();
                }
            }
            if (act) tma_wait_read<0>();
            if (act) op.store(bq, lane0 + j, sstride);
            return;
        }
    }

    // input prefetch (the input is 1/2^K of the traffic): float4 v = tid + j*NT of the tile
    float4 nxt[NVT];  // FM uses .x only
    auto fetch = [&](size_t tile) {
#pragma unroll
        for (int j = 0; j < NVT; j++) {
            const int v = tid + j * NT;
            if constexpr (FM) {
                const int plane = v % NL, t = v / NL;  // lane-fastest: 8 lanes of a frame are contiguous
                nxt[j].x = (v < NV && plane < nl) ? x[(tile * TI + t) * lanes + lane0 + plane] : 0.f;
            } else {
                const int plane = v / (TI / 4), pvec = v % (TI / 4);
                nxt[j] = (v < NV && plane < nl)
                             ? *reinterpret_cast<const float4 *>(x + (lane0 + plane) * n_in + tile * TI + 4 * pvec)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    if (ntiles) fetch(0);
    StateIO<K, 0, true>::run(sm, st, sstride, lane0, nl, tid);  // after the first fetch: one round trip for both

    for (size_t i = 0; i < ntiles; i++) {
        const int ob = (int)(i & 1);
#pragma unroll
        for (int j = 0; j < NVT; j++) {
            const int v = tid + j * NT;
            if constexpr (FM) {
                const int plane = v % NL, t = v / NL;
                if (v < NV) sm[off_u(K, 0) + plane * pitch(K, 0) + hist(0) + t] = nxt[j].x;
            } else {
                const int plane = v / (TI / 4), pvec = v % (TI / 4);
                if (v < NV) *reinterpret_cast<float4 *>(sm + off_u(K, 0) + plane * pitch(K, 0) + hist(0) + 4 * pvec) = nxt[j];
            }
        }
        if (i + 1 < ntiles) fetch(i + 1);
        // rows K-1 of the previous tile: last read in its final phase, next written by stage K-2
        if constexpr (K >= 2) {
            if (i > 0) carry_rows<K, K - 1>(sm, tid >> 5, tid & 31);
        }
        // the staging buffer about to be refilled must have been drained by its bulk stores
        // (bulk async-groups are per thread: every issuing thread waits for its own)
        if constexpr (!FM && !BQ) {
            if (tid < nl) tma_wait_read<1>();
        }
        if constexpr (FMT) {
            if (tid == 0) tma_wait_read<1>();
        }
        fir_sync();
        if constexpr (K >= 2) { StageRun<K, 0>::run(sm, tid, ob); fir_sync(); }
        if constexpr (K >= 3) { StageRun<K, 1>::run(sm, tid, ob); fir_sync(); }
        if constexpr (K >= 4) { StageRun<K, 2>::run(sm, tid, ob); fir_sync(); }
        if constexpr (K >= 5) { StageRun<K, 3>::run(sm, tid, ob); fir_sync(); }
        if constexpr (BQ) {
            if (i >= 2) nbar_sync(4 + ob, NTA);  // the biquad warp has stored tile i - 2 out of this buffer
        }
        StageRun<K, K - 1, FMT>::run(sm, tid, ob);  // -> staging (and carries rows K-2)
        if constexpr (BQ) {
            nbar_arrive(2 + ob, NTA);  // hand the staged tile to the biquad warp
            fir_sync();
            if constexpr (K == 1) {
                carry_rows<K, 0>(sm, tid >> 5, tid & 31);
                fir_sync();
            }
            continue;
        }
        if constexpr (!FM || FMT) fence_async_smem();  // writers make the staging rows visible to the async proxy
        __syncthreads();
        if constexpr (FMT) {
            using F = FmOut<FMT ? K : 4>;
            if (tid == 0) {
#pragma unroll
                for (int q = 0; q < F::NB; q++)  // box q: lanes lane0 + q * LPL .., frames i * FT ..
                    tma_store_2d(&ymap, smem_u32(sm + off_out(K)) + (uint32_t)(ob * F::BUF_LINES + q * F::LINES) * 128u,
                                 (int)((lane0 + q * F::LPL) * F::R), (int)(i * F::FT));
                tma_commit();
            }
        } else if constexpr (FM) {
            constexpr int R = 1 << K;
            constexpr int PF_ = R >= 4 ? 4 : 2;  // floats per piece: 16 bytes, or the 8-byte frame of x2
            const float *stg = sm + off_out(K) + ob * NL * OUT_PITCH;
            // thread -> (lane l = tid % NL, pieces q0, q0 + NT/NL, ...): piece q = output samples PF_*q .. of
            // lane l's tile; a pass advances by ADV samples = whole frames, so both pointers move by constants
            constexpr int QPT = NT / NL, ADV = PF_ * QPT;
            static_assert(NT % NL == 0 && ADV % R == 0, "a pass must advance every lane by whole frames");
            const int l = tid % NL, q0 = tid / NL;
            if (l < nl) {
                float *dst = y + ((i * TI + (PF_ * q0) / R) * lanes + lane0 + l) * R + (PF_ * q0) % R;
                const float *src = stg + l * OUT_PITCH + PF_ * q0;
                const size_t dstep = (size_t)ADV * lanes;
#pragma unroll 4
                for (int q = q0; q < TOUT / PF_; q += QPT, dst += dstep, src += ADV) {
                    if constexpr (PF_ == 4) *reinterpret_cast<float4 *>(dst) = lds128v(src);
                    else *reinterpret_cast<float2 *>(dst) = *reinterpret_cast<const float2 *>(src);
                }
            }
        } else if (tid < nl) {
            bulk_store_1d(y + (lane0 + tid) * n_out + i * TOUT, smem_u32(sm + off_out(K) + (ob * NL + tid) * OUT_PITCH),
                          TOUT * 4);
            tma_commit();
        }
        if constexpr (K == 1) {  // rows 0 are rewritten at the top of the next tile: carry them now
            carry_rows<K, 0>(sm, tid >> 5, tid & 31);
            __syncthreads();
        }
    }
    if constexpr (K >= 2) {
        carry_rows<K, K - 1>(sm, tid >> 5, tid & 31);
        fir_sync();
    }
    if constexpr (!FM && !BQ) {
        if (tid < nl) tma_wait_read<0>();
    }
    if constexpr (FMT) {
        if (tid == 0) tma_wait_read<0>();
    }
    StateIO<K, 0, false>::run(sm, st, sstride, lane0, nl, tid);
}

template <int K, bool FM, bool BQ = false>
static int launch(idsp_ctx *ctx, float *st, const float *x, float *y, size_t n_in, size_t ntiles, size_t lanes,
                  size_t sstride, const Df1Op<float, false>::Params &bq = Df1Op<float, false>::Params()) {
    unsigned grid = (unsigned)((lanes + NL - 1) / NL);
    CUtensorMap ymap;
    memset(&ymap, 0, sizeof(ymap));
    if constexpr (FM && !BQ && fm_out_rate(K)) {
        // y[frame][lane][R] as a 2-D tensor of (lanes * R) x frames words; box = one 128-byte line of lanes x the
        // tile's frames
        if (!getenv("IDSP_HBF_FM_LDGSTS") && lanes * (1ull << K) < (1ull << 31) && n_in < (1ull << 31) &&
            make_map_2d(&ymap, y, (uint64_t)lanes << K, (uint64_t)n_in, 32, FmOut<K>::FT, CU_TENSOR_MAP_SWIZZLE_128B)) {
            auto kern = hbf_int_fast_kernel<K, true, false, true>;
            IDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(K)));
            kern<<<grid, NT, smem_bytes(K), ctx->stream>>>(st, x, y, n_in, ntiles, lanes, sstride, bq, ymap);
            IDSP_KERNEL_FAMILY(ctx, "hbf tiled frame-major (tensor-map output)");
            IDSP_LAUNCHED(ctx);
            return IDSP_OK;
        }
    }
    auto kern = hbf_int_fast_kernel<K, FM, BQ, false>;
    IDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(K)));
    kern<<<grid, NT + (BQ ? 32 : 0), smem_bytes(K), ctx->stream>>>(st, x, y, n_in, ntiles, lanes, sstride, bq, ymap);
    IDSP_KERNEL_FAMILY(ctx, BQ ? "hbf tiled interpolator + fused biquad warp" : (FM ? "hbf tiled frame-major" : "hbf tiled lane-major"));
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

}  // namespace HFI_NS
}  // namespace idsp
