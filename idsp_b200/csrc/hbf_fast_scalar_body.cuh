// hbf_fast_scalar_body.cuh -- body of the tiled HBF /2^K decimation cascade, included once per tap set by
// hbf_fast_scalar.cuh with HFS_NS (namespace), HFS_TAPS (tap struct template) and HFS_M (tap count function)
// defined.  See hbf_fast_scalar.cuh for the description of the kernel.
namespace idsp {
namespace HFS_NS {


#ifndef HFS_NL
#define HFS_NL 8
#endif
constexpr int NL = HFS_NL;  // lanes per CTA (multiple of 8)
#ifndef HFS_NT
#define HFS_NT 128
#endif
constexpr int NT = HFS_NT;  // threads per CTA
#ifndef HFS_TT
#define HFS_TT 512
#endif
constexpr int TT = HFS_TT;  // raw input samples per lane per tile
constexpr int S = 2;      // raw ring depth

__host__ __device__ constexpr int up4(int v) { return (v + 3) & ~3; }
// pitch in floats: multiple of 4 with pitch/4 odd (conflict-free 16-byte accesses over 8 rows)
__host__ __device__ constexpr int oddpitch(int v) { return (up4(v) / 4) % 2 ? up4(v) : up4(v) + 4; }
__host__ __device__ constexpr int st_m(int K, int s) { return HFS_M(K - 1 - s); }
__host__ __device__ constexpr int st_n(int s) { return TT >> (s + 1); }  // outputs per lane per tile
#ifndef HFS_MINB
#define HFS_MINB 4
#endif
#ifndef HFS_R0
#define HFS_R0 16
#endif
#ifndef HFS_R0_K1
#define HFS_R0_K1 8
#endif
#ifndef HFS_BAL
#define HFS_BAL 0
#endif
// outputs per work item.  HFS_BAL: every stage has exactly NT items (no idle threads in the
// low-rate stages) at the price of more window loads per output there.
__host__ __device__ constexpr int st_r_bal(int s) {
    int r = NL * st_n(s) / NT;
    return r > 8 ? 8 : (r < 1 ? 1 : r);
}
__host__ __device__ constexpr int st_r(int s) {
    return s == 0 ? HFS_R0 : HFS_BAL ? st_r_bal(s) : (st_n(s) / 8 >= 8 ? 8 : (st_n(s) / 8 >= 4 ? st_n(s) / 8 : 4));
}
__host__ __device__ constexpr int raw_h(int K) { return up4(4 * st_m(K, 0) - 2); }
__host__ __device__ constexpr int raw_pitch(int K) { return oddpitch(raw_h(K) + TT); }
__host__ __device__ constexpr int he(int K, int s) { return up4(st_m(K, s) - 1); }
__host__ __device__ constexpr int ho(int K, int s) { return up4(2 * st_m(K, s) - 1); }
__host__ __device__ constexpr int pe(int K, int s) { return oddpitch(he(K, s) + st_n(s)); }
__host__ __device__ constexpr int po(int K, int s) { return oddpitch(ho(K, s) + st_n(s)); }
// float offsets inside dynamic shared memory
// the raw ring: S buffers of NL padded rows; the frame-major /16 tensor-map layout (FmTma: S x 4 boxes of
// TT/16 + 1 lines of 32 floats) lives in the same space
__host__ __device__ constexpr bool fm_tma_rate(int K) { return K >= 2 && K <= 5 && NL == 8 && TT == 512; }
__host__ __device__ constexpr int raw_floats(int K) {
    const int R = 1 << K, lpl = 32 / (R > 32 ? 32 : R), hf = (raw_h(K) + R - 1) / R;
    const int nb = NL / (lpl ? lpl : 1), br = TT / R + hf;
    const int rows = S * NL * raw_pitch(K), boxes = S * nb * (nb == 1 ? br : (br | 1)) * 32;
    return (fm_tma_rate(K) && boxes > rows) ? boxes : rows;
}
__host__ __device__ constexpr int off_e(int K, int s) {
    int o = raw_floats(K);
    for (int i = 1; i < s; i++) o += NL * (pe(K, i) + po(K, i));
    return o;
}
__host__ __device__ constexpr int off_o(int K, int s) { return off_e(K, s) + NL * pe(K, s); }
__host__ __device__ constexpr int smem_floats(int K) { return off_e(K, K); }
__host__ __device__ constexpr size_t smem_bytes(int K) { return (size_t)smem_floats(K) * 4 + S * 8; }
// ABI state word offset of stage s (highest-rate stage first): sum of 3M-2
__host__ __device__ constexpr int st_word(int K, int s) {
    int w = 0;
    for (int i = 0; i < s; i++) w += 3 * st_m(K, i) - 2;
    return w;
}

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// 128-bit shared loads as explicit PTX: the compiler must not split them into scalar LDS
// (only some components of a window are used, and scalar loads over rows of pitch 4*odd
// floats would be 4-way bank conflicted).
__device__ __forceinline__ float4 lds128(const float *p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(smem_u32(p)));
    return v;
}

// One item of the raw (interleaved) stage: outputs p0 .. p0+R-1 of lane row `row`
// (row[0..HR) = history, row[HR..] = tile).  Stream sample u[k] (k relative to the tile
// start) sits at row[HR + k]; the window starts at row[2*p0] (16-byte aligned).
template <int TI, int R> struct RawItem {
    static constexpr int M = HFS_TAPS<TI>::M;
    static constexpr int HR = up4(4 * M - 2);
    static constexpr int W = HR + 2 * R;
    __device__ __forceinline__ static void fir(const float (&w)[W], float (&y)[R]) {
#pragma unroll
        for (int q = 0; q < R; q++) {
            float acc = (w[2 * q + 1 + HR] + w[2 * q - 4 * M + 3 + HR]) * HFS_TAPS<TI>::c(0);
#pragma unroll
            for (int i = 1; i < M; i++)
                acc = acc + (w[2 * q - 2 * i + 1 + HR] + w[2 * q + 2 * i - 4 * M + 3 + HR]) * HFS_TAPS<TI>::c(i);
            y[q] = acc + w[2 * q - 2 * M + 2 + HR];
        }
    }
    __device__ __forceinline__ static void run(const float *row, int p0, float (&y)[R]) {
        float w[W];
        const float *src = row + 2 * p0;
#pragma unroll
        for (int j = 0; j < W / 4; j++) {
            float4 v = lds128(src + 4 * j);
            w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
        }
        fir(w, y);
    }
    // The same item on the frame-major tensor-map layout of /16 (see FmTma below): lane `lane`, outputs
    // 16 * j16 .. (R == 16, i.e. two frames of input behind one frame of history) of raw buffer `buf`.
    template <int K> __device__ __forceinline__ static void run_fm(uint32_t sm_base, int buf, int lane, int j16, float (&y)[R]);
};

// Frame-major input of the /16 cascade through the tensor-map unit.  The stream is x[frame][lane][16]: the two
// lanes of a pair are 128 contiguous bytes per frame, so a 2-D box of 32 floats x 33 frames (one frame of history
// in front of the 32 frames of a tile) is one lane pair's tile, and a CTA's 8 lanes are four such boxes per tile --
// four instructions instead of 1 056 16-byte LDGSTS copies.  The boxes land as [frame][lane pair][16] lines of 128
// bytes with the 128-byte swizzle, which XORs the 16-byte chunk index with bits 7-9 of the ABSOLUTE shared-memory
// address (measured, tools/tma_swizzle_probe.cu: destinations need 128-byte alignment only, rows before the start
// of the tensor are zero-filled).  The four boxes of a buffer are packed back to back (33 lines each), which
// staggers their swizzle phase: the eight lanes of a quarter-warp that read chunk c of the same frame hit eight
// different 16-byte slots -- conflict free without padding.  A window chunk is addressed as A ^ (c << 4) with a
// per-item, per-frame base A.
// General form (K = 2 ... 5): a frame of a lane is R = 2^K floats, LPL = 32 / R lanes share a 128-byte line, a CTA's
// NL lanes are NB = NL / LPL boxes of LINES = frames per tile + history frames; LINES is odd, so box q's swizzle
// phase is q lines ahead of box 0's.
template <int K> struct FmTma {
    static constexpr int R = 1 << K;               // floats per frame and lane
    static constexpr int LPL = 32 / R;             // lanes per 128-byte line
    static constexpr int NB = NL / LPL;            // boxes per raw buffer
    static constexpr int HF = (raw_h(K) + R - 1) / R;  // history frames in front of a tile
    static constexpr int BOX_ROWS = TT / R + HF;   // 128-byte lines a box delivers
    static constexpr int LINES = NB == 1 ? BOX_ROWS : (BOX_ROWS | 1);  // box pitch in lines (odd: see above)
    static constexpr int BUF_LINES = NB * LINES;   // per raw buffer
    static constexpr uint32_t BOX_BYTES = BOX_ROWS * 128;
    static_assert(LPL >= 1 && NL % LPL == 0 && (NB == 1 || LINES % 2 == 1), "staggered swizzle phases need an odd box height");
    // byte address of chunk 0 of `lane`'s frame in line `line` (absolute line index from the 1024-aligned base)
    __device__ __forceinline__ static uint32_t line_base(uint32_t sm_base, int line, int lane) {
        return sm_base + (uint32_t)line * 128u + (uint32_t)(((((lane % LPL) * (R / 4)) ^ line) & 7) << 4);
    }
    __device__ __forceinline__ static int line_of(int buf, int lane, int row) {
        return buf * BUF_LINES + (lane / LPL) * LINES + row;
    }
    // address of one float: lane, box row (0 = first history frame) and sample 0..R-1 inside the frame
    __device__ __forceinline__ static uint32_t word(uint32_t sm_base, int buf, int lane, int row, int smp) {
        return (line_base(sm_base, line_of(buf, lane, row), lane) ^ (uint32_t)((smp >> 2) << 4)) + (uint32_t)(smp & 3) * 4u;
    }
};
__device__ __forceinline__ float4 lds128a(uint32_t addr) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}
template <int TI, int R>
template <int K>
__device__ __forceinline__ void RawItem<TI, R>::run_fm(uint32_t sm_base, int buf, int lane, int j16, float (&y)[R]) {
    using F = FmTma<K>;
    static_assert(R == 16 && 32 % F::R == 0, "an item is 32 input samples = a whole number of frames");
    static_assert(HR % 4 == 0 && HR <= F::HF * F::R, "the raw history fits the frames in front of the tile");
    constexpr int NR = F::HF + 32 / F::R;  // box rows an item's window touches
    float w[W];
    const int line0 = F::line_of(buf, lane, j16 * (32 / F::R));  // first history row of the item
    uint32_t A[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) A[r] = F::line_base(sm_base, line0 + r, lane);
#pragma unroll
    for (int j = 0; j < W / 4; j++) {
        const int smp = F::HF * F::R - HR + 4 * j;  // sample index counted from the start of row line0
        float4 v = lds128a(A[smp / F::R] ^ (uint32_t)(((smp % F::R) / 4) << 4));
        w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
    fir(w, y);
}

// One item of a de-interleaved stage: erow = [HE hist | n new], orow = [HO hist | n new].
// The item spans R outputs starting at p0 (p0 % 4 == 0 keeps the LDS.128 aligned); this call
// evaluates outputs Q0 .. Q0+QN-1 of it and only loads the part of the window they need.
template <int TI, int R, int Q0 = 0, int QN = R> struct SplitItem {
    static constexpr int M = HFS_TAPS<TI>::M;
    static constexpr int LEN = 2 * M - 1;
    static constexpr int HE = up4(M - 1), HO = up4(LEN);
    static constexpr int RE = HE - (M - 1), RO = HO - LEN;
    static constexpr int JO0 = (RO + Q0) / 4, JO1 = (RO + Q0 + QN + 2 * M - 2) / 4 + 1;
    static constexpr int JE0 = (RE + Q0) / 4, JE1 = (RE + Q0 + QN - 1) / 4 + 1;
    __device__ __forceinline__ static void run(const float *erow, const float *orow, int p0, float (&y)[QN]) {
        float wo[4 * JO1], we[4 * JE1];
#pragma unroll
        for (int j = JO0; j < JO1; j++) {
            float4 v = lds128(orow + p0 + 4 * j);
            wo[4 * j] = v.x; wo[4 * j + 1] = v.y; wo[4 * j + 2] = v.z; wo[4 * j + 3] = v.w;
        }
#pragma unroll
        for (int j = JE0; j < JE1; j++) {
            float4 v = lds128(erow + p0 + 4 * j);
            we[4 * j] = v.x; we[4 * j + 1] = v.y; we[4 * j + 2] = v.z; we[4 * j + 3] = v.w;
        }
#pragma unroll
        for (int q = Q0; q < Q0 + QN; q++) {
            float acc = (wo[RO + q + 2 * M - 1] + wo[RO + q]) * HFS_TAPS<TI>::c(0);
#pragma unroll
            for (int i = 1; i < M; i++)
                acc = acc + (wo[RO + q + 2 * M - 1 - i] + wo[RO + q + i]) * HFS_TAPS<TI>::c(i);
            y[q - Q0] = acc + we[RE + q];
        }
    }
};

// scatter R consecutive outputs (p0 multiple of R, R in {4,8}) into the next stage's E/O rows
template <int R>
__device__ __forceinline__ void put_split(float *erow_new, float *orow_new, int p0, const float (&y)[R]) {
    // erow_new / orow_new already point at the first NEW element (past the history)
    if constexpr (R >= 8) {
#pragma unroll
        for (int j = 0; j < R / 8; j++) {
            reinterpret_cast<float4 *>(erow_new + p0 / 2)[j] = make_float4(y[8 * j], y[8 * j + 2], y[8 * j + 4], y[8 * j + 6]);
            reinterpret_cast<float4 *>(orow_new + p0 / 2)[j] = make_float4(y[8 * j + 1], y[8 * j + 3], y[8 * j + 5], y[8 * j + 7]);
        }
    } else if constexpr (R == 4) {
        *reinterpret_cast<float2 *>(erow_new + p0 / 2) = make_float2(y[0], y[2]);
        *reinterpret_cast<float2 *>(orow_new + p0 / 2) = make_float2(y[1], y[3]);
    } else if constexpr (R == 2) {
        erow_new[p0 / 2] = y[0];
        orow_new[p0 / 2] = y[1];
    } else {
        ((p0 & 1) ? orow_new : erow_new)[p0 / 2] = y[0];
    }
}

// Move the tails of the [hist | n new] rows E_s / O_s of all NL lanes to their heads (the
// reference's copy_within, src/hbf.rs:183-184).  Executed by `nw` warps (`wsel` = index of
// this warp among them); one 16-byte piece per thread, a whole row inside one warp so that
// reading everything before writing anything (head and tail overlap when hist > n) only
// needs a __syncwarp.
template <int K, int s>
__device__ __forceinline__ void carry_rows(float *sm, int wsel, int nw, int lid) {
    constexpr int CE = he(K, s) / 4, CO = ho(K, s) / 4, C = CE + CO;
    static_assert(C <= 32, "row history too long for one warp");
    // a lane's C pieces sit in a group of CP = 2^k >= C consecutive threads (shift / mask; threads j >= C of a
    // group idle): LP lanes per warp pass.  (The earlier `lid / C`, `lid % C` form with LP = 32 / C was evaluated
    // with `act` false for every thread in ONE instantiation -- HBF_TAPS_98, /32, rows 2, C = 3, LP = 10 > NL --
    // although the same constants work one stage earlier: device printf showed sub = 0, lane = 0, act = 0.  No
    // race, no memory error (compute-sanitizer clean); this form does not trigger it and is cheaper.)
    constexpr int CP = C <= 1 ? 1 : C <= 2 ? 2 : C <= 4 ? 4 : C <= 8 ? 8 : C <= 16 ? 16 : 32;
    constexpr int LP = 32 / CP;
    const int sub = lid / CP, j = lid % CP;
    const bool odd = j >= CE;
    const int pitch = odd ? po(K, s) : pe(K, s);
    const int joff = odd ? off_o(K, s) + 4 * (j - CE) : off_e(K, s) + 4 * j;
    for (int l0 = wsel * LP; l0 < NL; l0 += nw * LP) {  // l0 is warp-uniform
        const int lane = l0 + sub;
        const bool act = j < C && lane < NL;
        float *row = sm + joff + lane * pitch;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) v = lds128(row + st_n(s));
        __syncwarp();
        if (act) *reinterpret_cast<float4 *>(row) = v;
        __syncwarp();
    }
}

template <int K, int s> struct StageRun {
    static constexpr int TI = K - 1 - s;
    static constexpr int R = st_r(s);          // outputs per work item
    static constexpr int RA = R < 4 ? 4 : R;   // outputs per aligned window (p0 % 4 == 0)
    static constexpr int NSUB = RA / R;        // work items sharing one window (warp-uniform split)
    static constexpr int WIN = NL * st_n(s) / RA;
    static_assert(NSUB == 1 || WIN % 32 == 0, "sub-item index must be warp-uniform");

    template <int Q0>
    // y is lane-major (row stride `ystride`) when ylanes == 0, else frame-major with `ylanes` lanes
    __device__ __forceinline__ static void item(float *sm, int lane, int p0, int nl, float *y, size_t ystride,
                                                size_t yoff, size_t lane0, size_t ylanes) {
        const float *E = sm + off_e(K, s);
        const float *O = sm + off_o(K, s);
        float out[R];
        SplitItem<TI, RA, Q0, R>::run(E + lane * pe(K, s), O + lane * po(K, s), p0, out);
        if constexpr (s == K - 1) {
            if (lane < nl && ylanes) {
                float *dst = y + (yoff + p0 + Q0) * ylanes + lane0 + lane;
#pragma unroll
                for (int j = 0; j < R; j++) dst[(size_t)j * ylanes] = out[j];
            } else if (lane < nl) {
                float *dst = y + (lane0 + lane) * ystride + yoff + p0 + Q0;
                if (R >= 4 && (((uintptr_t)dst) & 15) == 0) {
#pragma unroll
                    for (int j = 0; j < R / 4; j++)
                        reinterpret_cast<float4 *>(dst)[j] =
                            make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
                } else if (R == 2 && (((uintptr_t)dst) & 7) == 0) {
                    *reinterpret_cast<float2 *>(dst) = make_float2(out[0], out[R - 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < R; j++) dst[j] = out[j];
                }
            }
        } else {
            float *En = sm + off_e(K, s + 1) + lane * pe(K, s + 1) + he(K, s + 1);
            float *On = sm + off_o(K, s + 1) + lane * po(K, s + 1) + ho(K, s + 1);
            put_split<R>(En, On, p0 + Q0, out);
        }
    }

    // Warps that hold items of this stage: low-rate stages have fewer than NT items, and
    // odd stages take the upper warps so that, over the CTAs of an SM, every scheduler gets
    // work.  The other warps move the previous stage's row tails to the heads meanwhile.
    static constexpr int ITEMS = WIN * NSUB;
    static constexpr int NW = NT / 32;
    static constexpr int NWA = ITEMS >= NT ? NW : (ITEMS + 31) / 32;
    static constexpr int W0 = (s & 1) ? NW - NWA : 0;

    static constexpr int NIDLE = 32 * (NW - NWA);  // threads without items in this phase

    // runs stage s (1 <= s <= K-1) for one tile; `idle(gt, NIDLE)` is extra work for the
    // threads of the warps that hold no items (after they carried rows s-1)
    template <class F>
    __device__ __forceinline__ static void run(float *sm, int tid, int nl, float *y, size_t ystride,
                                               size_t yoff, size_t lane0, size_t ylanes, F &&idle) {
        const int vt = tid - 32 * W0;
        if (vt >= 0 && vt < 32 * NWA) {
            for (int idx = vt; idx < ITEMS; idx += 32 * NWA) {
                const int w = idx % WIN, sub = idx / WIN;
                const int lane = w % NL, p0 = (w / NL) * RA;
                if constexpr (NSUB == 1) {
                    item<0>(sm, lane, p0, nl, y, ystride, yoff, lane0, ylanes);
                } else if constexpr (NSUB == 2) {
                    if (sub == 0) item<0>(sm, lane, p0, nl, y, ystride, yoff, lane0, ylanes);
                    else item<R>(sm, lane, p0, nl, y, ystride, yoff, lane0, ylanes);
                } else {
                    if (sub == 0) item<0>(sm, lane, p0, nl, y, ystride, yoff, lane0, ylanes);
                    else if (sub == 1) item<R>(sm, lane, p0, nl, y, ystride, yoff, lane0, ylanes);
                    else if (sub == 2) item<2 * R>(sm, lane, p0, nl, y, ystride, yoff, lane0, ylanes);
                    else item<3 * R>(sm, lane, p0, nl, y, ystride, yoff, lane0, ylanes);
                }
            }
        }
        if constexpr (s >= 2) {  // rows s-1 were consumed in the previous phase
            const int warp = tid >> 5;
            if constexpr (NWA < NW) {
                if (vt < 0 || vt >= 32 * NWA) {
                    const int iw = vt < 0 ? warp : warp - NWA;  // index among the idle warps
                    carry_rows<K, s - 1>(sm, iw, NW - NWA, tid & 31);
                    idle(32 * iw + (tid & 31), NIDLE);
                }
            } else {
                carry_rows<K, s - 1>(sm, warp, NW, tid & 31);
            }
        }
    }
    __device__ __forceinline__ static void run(float *sm, int tid, int nl, float *y, size_t ystride,
                                               size_t yoff, size_t lane0, size_t ylanes) {
        run(sm, tid, nl, y, ystride, yoff, lane0, ylanes, [](int, int) {});
    }
};

// barrier among the `count` threads of the idle warps of a phase (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void bar_idle(int count) { asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory"); }

// ABI state <-> shared-memory histories (see header comment of include/idsp_b200.h)
// WHICH: 0 = every stage, 1 = stage 0 only, 2 = every stage but stage 0.  FMT: stage 0's raw history lives in the
// tensor-map layout (FmTma) of raw buffer `rawbuf`, in row 0 (head) on entry and in the tile's last row on exit.
template <int K, int s, bool LOAD, bool FMT = false, int WHICH = 0> struct StateIO {
    // raw history lives at row[roff .. roff+HR): roff = 0 (head) on entry, TT (tail of the
    // last tile) on exit
    __device__ __forceinline__ static void run(float *sm, float *st, size_t sstride, size_t lane0, int nl,
                                               int tid, int rawbuf, int roff) {
        if constexpr (s < K) {
            constexpr int M = st_m(K, s);
            constexpr int LEN = 2 * M - 1;
            constexpr int WORDS = 3 * M - 2;
            constexpr bool SKIP = (WHICH == 1 && s != 0) || (WHICH == 2 && s == 0);
            float *stw = st + (size_t)st_word(K, s) * sstride + lane0;
            // shared-memory slot of state word w of lane `lane`
            auto slot = [&](int lane, int w) -> float * {
                if constexpr (s == 0) {
                    constexpr int HR = raw_h(K);
                    const int hs = w < M - 1 ? (HR - 2 * M + 2 + 2 * w) : (HR - 4 * M + 3 + 2 * (w - (M - 1)));
                    if constexpr (FMT) {
                        using F = FmTma<K>;
                        const int sp = roff + F::HF * F::R - HR + hs;  // sample counted from the first history row
                        const uint32_t a = F::word(smem_u32(sm), rawbuf, lane, sp / F::R, sp % F::R);
                        return sm + (a - smem_u32(sm)) / 4;
                    } else {
                        return sm + (rawbuf * NL + lane) * raw_pitch(K) + roff + hs;
                    }
                } else {
                    return w < M - 1 ? sm + off_e(K, s) + lane * pe(K, s) + (he(K, s) - (M - 1) + w)
                                     : sm + off_o(K, s) + lane * po(K, s) + (ho(K, s) - LEN + (w - (M - 1)));
                }
            };
            if constexpr (SKIP) {
            } else if constexpr (LOAD) {
                // all the loads of a thread are issued before the first store (read-only path: the compiler may
                // hoist them over the shared-memory stores of the previous stage too), so that entering a call
                // costs one global-memory round trip, not one per state word: short calls of a streaming
                // caller are dominated by this prologue otherwise (13 % of a 16-tile call)
                constexpr int IT = (WORDS * NL + NT - 1) / NT;
                float v[IT];
#pragma unroll
                for (int it = 0; it < IT; it++) {
                    const int idx = tid + it * NT, lane = idx % NL, w = idx / NL;
                    v[it] = (idx < WORDS * NL && lane < nl) ? __ldg(stw + (size_t)w * sstride + lane) : 0.f;
                }
#pragma unroll
                for (int it = 0; it < IT; it++) {
                    const int idx = tid + it * NT, lane = idx % NL, w = idx / NL;
                    if (idx < WORDS * NL && lane < nl) *slot(lane, w) = v[it];
                }
            } else {
                for (int idx = tid; idx < WORDS * NL; idx += NT) {
                    const int lane = idx % NL, w = idx / NL;
                    if (lane >= nl) continue;
                    stw[(size_t)w * sstride + lane] = *slot(lane, w);
                }
            }
            StateIO<K, s + 1, LOAD, FMT, WHICH>::run(sm, st, sstride, lane0, nl, tid, rawbuf, roff);
        }
    }
};

// 16-byte asynchronous copy global -> shared (LDGSTS, L1 bypassed) and its mbarrier hook
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {  // arrives when this thread's copies landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// FM = false: x, y lane-major (x rows of n_in floats).  FM = true: frame-major,
// x[t][lane][2^K], y[t][lane]: the per-lane rows of a tile are gathered with 16-byte (/2: 8-byte) LDGSTS
// copies (8 lanes x 4 pieces of one frame per warp instruction = 512 contiguous bytes of HBM,
// 8 different shared-memory rows per quarter-warp), every thread arriving on the tile's mbarrier.
// FMT (frame-major /16 only): the input tiles arrive as tensor-map boxes (FmTma) instead of LDGSTS pieces.
template <int K, bool FM, bool FMT = false>
__global__ void __launch_bounds__(NT, HFS_MINB)
hbf_dec_fast_kernel(float *st, const float *x, float *y, size_t n_out, size_t ntiles, size_t lanes,
                    size_t sstride, const __grid_constant__ CUtensorMap xmap) {
    static_assert(!FMT || (FM && fm_tma_rate(K) && S * FmTma<FMT ? K : 4>::BUF_LINES * 32 <= raw_floats(K)),
                  "tensor-map input: frame-major /4 ... /32, 8 lanes x 512 samples, inside the raw ring");
    const size_t ylanes = FM ? lanes : 0;
    constexpr int TI0 = K - 1;
    // /2 runs the 23-tap stage on the raw stream: 16 outputs per item need a 124-float window and spill
    // (96 bytes of stack, reloaded on the refill path); 8 outputs per item fit in registers
    constexpr int R0 = K == 1 ? HFS_R0_K1 : st_r(0);
    constexpr int HR = raw_h(K);
    constexpr int PR = raw_pitch(K);
    constexpr int TO = TT >> K;
    extern __shared__ __align__(1024) float sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + smem_floats(K));
    const int tid = threadIdx.x;
    const size_t lane0 = (size_t)blockIdx.x * NL;
    const int nl = (int)((lanes - lane0) < (size_t)NL ? (lanes - lane0) : (size_t)NL);
    const size_t n_in = n_out << K;  // row stride of x in floats

    // zero everything once (unused history slots / absent lanes must hold finite garbage-free data)
    for (int i = tid; i < smem_floats(K); i += NT) sm[i] = 0.f;
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < S; b++) mbar_init(smem_u32(&bars[b]), (FM && !FMT) ? NT : 1);
        mbar_fence_init();
    }
    __syncthreads();
    // (FMT: the raw history of tile 0 is scattered after the tile has landed -- its box covers the history row)
    StateIO<K, 0, true, FMT, FMT ? 2 : 0>::run(sm, st, sstride, lane0, nl, tid, 0, 0);
    // generic-proxy writes above (zero fill) precede async-proxy (TMA) writes to the same rows
    fence_async_smem();
    __syncthreads();

    // Tile t >= 1 is fetched together with the HR samples in front of it (they are in L2 from
    // the previous tile), so the raw history never has to be copied between ring buffers;
    // tile 0 takes its history from the ABI state (scattered into buffer 0 above).
    // Executed by every warp: lane 0 of warp w issues the rows of lanes w*LPW .. (NL/NW rows each), so
    // no warp is held up by a serial chain of NL bulk copies (complete_tx may precede the
    // expect_tx of thread 0: the phase cannot complete before that arrival).
    auto issue = [&](size_t tile) {
        const int b = (int)(tile % S);
        const uint32_t bar = smem_u32(&bars[b]);
        const uint32_t hist = tile ? HR : 0;
        if constexpr (FMT) {
            using F = FmTma<FMT ? K : 4>;
            if (tid == 0) {
                mbar_expect_tx(bar, F::NB * F::BOX_BYTES);
#pragma unroll
                for (int q = 0; q < F::NB; q++)  // box q: lanes lane0 + q * LPL .., frames tile * (TT / R) - HF ..
                    tma_load_2d(smem_u32(sm) + (uint32_t)(b * F::BUF_LINES + q * F::LINES) * 128u, &xmap,
                                (int)((lane0 + q * F::LPL) * F::R), (int)(tile * (TT / F::R)) - F::HF, bar);
            }
        } else if constexpr (FM) {
            constexpr int R = 1 << K;            // floats per frame and lane
            constexpr int PF_ = R >= 4 ? 4 : 2;  // floats per piece: 16 bytes, or the whole 8-byte frame of /2
            // piece q = PF_ consecutive stream samples of one lane, counted from the start of the history.
            // Thread -> (lane l = tid % NL, pieces q0, q0 + NT/NL, ...): the lane is fixed per thread and a pass
            // advances the stream by ADV = PF_ * NT / NL samples = a whole number of frames, so source and
            // destination move by constant strides (one 64-bit and one 32-bit add per piece instead of the
            // divisions / multiplications of the general index map).
            constexpr int QPT = NT / NL;  // pieces of one lane per pass
            constexpr int ADV = PF_ * QPT;
            static_assert(NT % NL == 0 && ADV % R == 0, "a pass must advance every lane by whole frames");
            const int nq = (int)((TT + hist) / PF_);  // pieces per lane
            const size_t s0 = tile * TT - hist;     // stream position of piece 0
            const int l = tid % NL, q0 = tid / NL;
            if (l < nl) {
                const size_t sp0 = s0 + PF_ * (size_t)q0;
                const float *src = x + ((sp0 / R) * lanes + lane0 + l) * R + (sp0 % R);
                uint32_t dst = smem_u32(sm + b * NL * PR + HR - hist + l * PR + PF_ * q0);
                const size_t sstep = (size_t)ADV * lanes;
#pragma unroll 3
                for (int q = q0; q < nq; q += QPT, src += sstep, dst += ADV * 4) {
                    if constexpr (PF_ == 4) cp_async16(dst, src);
                    else cp_async8(dst, src);
                }
            }
            cp_async_arrive(bar);
        } else if ((tid & 31) == 0) {
            constexpr int LPW = (NL + NT / 32 - 1) / (NT / 32);
            if (tid == 0) mbar_expect_tx(bar, (uint32_t)(nl * (TT + hist) * 4));
#pragma unroll
            for (int j = 0; j < LPW; j++) {
                const int l = (tid >> 5) * LPW + j;
                if (l < nl)
                    bulk_load_1d(smem_u32(sm + (b * NL + l) * PR + HR - hist),
                                 x + (lane0 + l) * n_in + tile * TT - hist, (TT + hist) * 4, bar);
            }
        }
    };
#pragma unroll
    for (int b = 0; b < S; b++)
        if ((size_t)b < ntiles) issue(b);
    if constexpr (FMT) {
        // tile 0 has landed (its history row zero-filled: frame -1 does not exist): now the state goes there
        if (ntiles) mbar_wait(smem_u32(&bars[0]), 0);
        StateIO<K, 0, true, true, 1>::run(sm, st, sstride, lane0, nl, tid, 0, 0);
        __syncthreads();
    }
    // steady state (tile >= S >= 1, lane-major): source / destination of this warp's rows are kept in
    // registers, so a refill is one multiply-add per pointer plus the copy itself
    constexpr int LPW_ = (NL + NT / 32 - 1) / (NT / 32);
    const float *isrc[LPW_];
    uint32_t idst[LPW_];
#pragma unroll
    for (int j = 0; j < LPW_; j++) {
        const int l = (tid >> 5) * LPW_ + j;
        isrc[j] = x + (lane0 + (l < nl ? l : 0)) * n_in - HR;
        idst[j] = smem_u32(sm + l * PR);
    }
    auto refill = [&](size_t tile) {
        if constexpr (FM) {
            issue(tile);
        } else if ((tid & 31) == 0) {
            const uint32_t b = (uint32_t)(tile % S);
            const uint32_t bar = smem_u32(&bars[b]);
            if (tid == 0) mbar_expect_tx(bar, (uint32_t)(nl * (TT + HR) * 4));
#pragma unroll
            for (int j = 0; j < LPW_; j++)
                if ((tid >> 5) * LPW_ + j < nl)
                    bulk_load_1d(idst[j] + b * (uint32_t)(NL * PR * 4), isrc[j] + tile * TT, (TT + HR) * 4, bar);
        }
    };

    // stage 0 of tile `t`: items [c0, c1) spread over the `G` threads of a group (gt = index in it)
    constexpr int ITEMS0 = NL * st_n(0) / R0;
    auto stage0 = [&](size_t t, int c0, int c1, int gt, int G) {
        const int b = (int)(t % S);
        mbar_wait(smem_u32(&bars[b]), (uint32_t)((t / S) & 1));
        const float *raw = sm + b * NL * PR;
        for (int idx = c0 + gt; idx < c1; idx += G) {
            const int lane = idx % NL, p0 = (idx / NL) * R0;
            float out[R0];
            if constexpr (FMT) RawItem<TI0, R0>::template run_fm<FMT ? K : 4>(smem_u32(sm), b, lane, idx / NL, out);
            else RawItem<TI0, R0>::run(raw + lane * PR, p0, out);
            if constexpr (K == 1) {
                if (lane < nl && FM) {
                    float *dst = y + (t * TO + p0) * lanes + lane0 + lane;
#pragma unroll
                    for (int j = 0; j < R0; j++) dst[(size_t)j * lanes] = out[j];
                } else if (lane < nl) {
                    float *dst = y + (lane0 + lane) * n_out + t * TO + p0;
                    if ((((uintptr_t)dst) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < R0 / 4; j++)
                            reinterpret_cast<float4 *>(dst)[j] =
                                make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < R0; j++) dst[j] = out[j];
                    }
                }
            } else {
                float *En = sm + off_e(K, 1) + lane * pe(K, 1) + he(K, 1);
                float *On = sm + off_o(K, 1) + lane * po(K, 1) + ho(K, 1);
                put_split<R0>(En, On, p0, out);
            }
        }
    };

    // K >= 4: the warps without items in the low-rate phases 2 and 3 run stage 0 of the NEXT tile
    // there (its raw data is already in the ring), so in steady state there is no stage-0 phase,
    // all warps are busy in every phase and a tile costs K-1 barriers.
    constexpr bool PF = K >= 4 && StageRun<K, 2>::NIDLE > 0 && StageRun<K, 3>::NIDLE > 0 &&
                        StageRun<K, 2>::NIDLE + StageRun<K, 3>::NIDLE >= ITEMS0;
    if constexpr (PF) {
        stage0(0, 0, ITEMS0, tid, NT);
        __syncthreads();
        if ((size_t)S < ntiles) refill(S);
    }
    for (size_t i = 0; i < ntiles; i++) {
        if constexpr (!PF) {
            // ---- phase 0: raw interleaved -> E_1 / O_1 (or -> y when K == 1)
            stage0(i, 0, ITEMS0, tid, NT);
            // rows K-1 of the previous tile (last read in its final phase, next written in phase K-2 >= 1)
            if constexpr (K >= 3) {
                if (i > 0) carry_rows<K, K - 1>(sm, tid >> 5, NT / 32, tid & 31);
            }
            __syncthreads();
            // ---- raw buffer b is free again: refill it
            if (i + S < ntiles) refill(i + S);
            // ---- phases 1 .. K-1 (phase s also carries rows s-1)
            if constexpr (K >= 2) { StageRun<K, 1>::run(sm, tid, nl, y, n_out, i * TO, lane0, ylanes); __syncthreads(); }
            if constexpr (K >= 3) { StageRun<K, 2>::run(sm, tid, nl, y, n_out, i * TO, lane0, ylanes); __syncthreads(); }
            if constexpr (K >= 4) { StageRun<K, 3>::run(sm, tid, nl, y, n_out, i * TO, lane0, ylanes); __syncthreads(); }
            if constexpr (K >= 5) { StageRun<K, 4>::run(sm, tid, nl, y, n_out, i * TO, lane0, ylanes); __syncthreads(); }
            if constexpr (K == 2) {  // rows 1 are written again in the very next phase: carry them now
                carry_rows<K, 1>(sm, tid >> 5, NT / 32, tid & 31);
                __syncthreads();
            }
        } else {
            constexpr int H0 = StageRun<K, 2>::NIDLE < ITEMS0 ? StageRun<K, 2>::NIDLE : ITEMS0;  // items done in phase 2
            const bool more = i + 1 < ntiles;
            // ---- phase 1: stage 1, then rows K-1 of the previous tile (next written in phase K-2 >= 2)
            StageRun<K, 1>::run(sm, tid, nl, y, n_out, i * TO, lane0, ylanes);
            if (i > 0) carry_rows<K, K - 1>(sm, tid >> 5, NT / 32, tid & 31);
            __syncthreads();
            // ---- phase 2: stage 2 | carry rows 1, then first part of stage 0 of tile i+1 (writes rows 1)
            StageRun<K, 2>::run(sm, tid, nl, y, n_out, i * TO, lane0, ylanes, [&](int gt, int G) {
                bar_idle(G);  // every row-1 tail has been carried before any is overwritten
                if (more) stage0(i + 1, 0, H0, gt, G);
            });
            __syncthreads();
            // ---- phase 3: stage 3 | carry rows 2, rest of stage 0 of tile i+1
            StageRun<K, 3>::run(sm, tid, nl, y, n_out, i * TO, lane0, ylanes, [&](int gt, int G) {
                if (more) stage0(i + 1, H0, ITEMS0, gt, G);
            });
            __syncthreads();
            if constexpr (K >= 5) { StageRun<K, 4>::run(sm, tid, nl, y, n_out, i * TO, lane0, ylanes); __syncthreads(); }
            // ---- the raw buffer of tile i+1 is free again: refill it
            if (i + 1 + S < ntiles) refill(i + 1 + S);
        }
    }
    if constexpr (K >= 3) {
        carry_rows<K, K - 1>(sm, tid >> 5, NT / 32, tid & 31);
        __syncthreads();
    }
    // the raw history of the stream is the tail of the last tile's buffer
    StateIO<K, 0, false, FMT>::run(sm, st, sstride, lane0, nl, tid, (int)((ntiles - 1) % S), TT);
}

#ifndef HFS_FM_TMA
#define HFS_FM_TMA 1
#endif
template <int K, bool FM>
static int launch(idsp_ctx *ctx, float *st, const float *x, float *y, size_t n_out, size_t ntiles,
                  size_t lanes, size_t sstride) {
    size_t smem = smem_bytes(K);
#ifdef IDSP_TUNE
    // occupancy experiment: pad the dynamic shared memory so that fewer CTAs fit on an SM (is the kernel
    // bound by issue slots or by latency?)
    if (getenv("IDSP_HBF_EXTRA_SMEM")) smem += (size_t)atoi(getenv("IDSP_HBF_EXTRA_SMEM"));
#endif
    unsigned grid = (unsigned)((lanes + NL - 1) / NL);
    CUtensorMap xmap;
    memset(&xmap, 0, sizeof(xmap));
    if constexpr (FM && fm_tma_rate(K) && HFS_FM_TMA) {
        // x[frame][lane][R] as a 2-D tensor of (lanes * R) x frames words; box = one 128-byte line of lanes x
        // (history + tile) frames
        const bool tune_off = getenv("IDSP_HBF_FM_LDGSTS") != nullptr;  // A/B switch: the LDGSTS gather
        if (!tune_off && lanes * (1ull << K) < (1ull << 31) && n_out < (1ull << 31) &&
            make_map_2d(&xmap, x, (uint64_t)lanes << K, (uint64_t)n_out, 32, FmTma<K>::BOX_ROWS, CU_TENSOR_MAP_SWIZZLE_128B)) {
            auto kern = hbf_dec_fast_kernel<K, true, true>;
            IDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, NT, smem, ctx->stream>>>(st, x, y, n_out, ntiles, lanes, sstride, xmap);
            IDSP_KERNEL_FAMILY(ctx, "hbf tiled frame-major (tensor-map input)");
            IDSP_LAUNCHED(ctx);
            return IDSP_OK;
        }
    }
    auto kern = hbf_dec_fast_kernel<K, FM, false>;
    IDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, ctx->stream>>>(st, x, y, n_out, ntiles, lanes, sstride, xmap);
    IDSP_KERNEL_FAMILY(ctx, FM ? "hbf tiled frame-major" : "hbf tiled lane-major");
    IDSP_LAUNCHED(ctx);
    return IDSP_OK;
}

}  // namespace HFS_NS
}  // namespace idsp
