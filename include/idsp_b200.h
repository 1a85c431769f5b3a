/*
 * idsp_b200.h -- C ABI of the B200 multi-lane sample-processing engine.
 *
 * Drop-in boundary for the filter hot path of quartiq/idsp: every entry point
 * replaces the inner loop of one reference `SplitProcess::block()` /
 * `Lanes::process_view()` / PyO3 function and cites it (paths relative to the
 * reference tree).  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *  - return value: 0 = ok, <0 = idsp_status_t error; idsp_b200_last_error()
 *    returns a thread-local message.  Nothing aborts or throws across the ABI.
 *    (The reference hot path is infallible; length mismatches are caller
 *    preconditions there -- dsp-process/src/process.rs:42-45,121-123 -- here they
 *    are IDSP_EINVAL.)
 *  - one idsp_ctx = one device + one CUDA stream; calls are asynchronous on that
 *    stream; idsp_b200_sync() waits.  A ctx is not thread-safe (mirrors `&mut`
 *    state exclusivity); different ctxs are independent.
 *  - unless a function name ends in `_host`, sample and state pointers are
 *    DEVICE pointers owned by the caller; coefficient pointers (`ba`, `clamp`,
 *    `k`, `taps`) are always small HOST arrays copied at launch.
 *    `_host` variants take host pointers for samples and state, stream them
 *    through the device in chunks (H2D / compute / D2H overlapped) and return
 *    after the results are in host memory.
 *  - layout: IDSP_FRAME_MAJOR  flat[t*lanes + l]  = `[[T; N]]` frames
 *            (dsp-process/src/view.rs:106-131, compose.rs:468-476)
 *            IDSP_LANE_MAJOR   flat[l*frames + t] = View<LaneMajor>
 *            (dsp-process/src/view.rs:176-225, compose.rs:478-494)
 *    Multi-sample frames (HBF: X = [f32; R]) keep R innermost:
 *    frame-major flat[(t*lanes + l)*R + r], lane-major flat[(l*frames + t)*R + r].
 *  - state is caller-owned, persists across calls (= streaming; the reference
 *    `block()` may be called repeatedly on the same state) and is SoA over
 *    lanes: state[word*lanes + lane].  Word order is the reference struct's
 *    field order and is given per function.  Zero-initialised state = the
 *    reference's `Default`.  x and y may alias exactly (in-place) wherever the
 *    reference implements SplitInplace.
 *  - integer arithmetic wraps like Rust release builds; f32/f64 arithmetic is
 *    IEEE round-to-nearest per operation, never fused, denormals preserved:
 *    results are bit-identical to the reference.
 */
#ifndef IDSP_B200_H
#define IDSP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDSP_B200_VERSION 100

typedef struct idsp_ctx idsp_ctx;

typedef enum {
    IDSP_OK = 0,
    IDSP_EINVAL = -1,   /* bad argument (null pointer, F out of range, size mismatch) */
    IDSP_ECUDA = -2,    /* CUDA runtime/driver error, see idsp_b200_last_error() */
    IDSP_ENOMEM = -3,   /* allocation failed */
    IDSP_ENODEV = -4,   /* no usable sm_100 device */
    IDSP_ENCCL = -5     /* NCCL could not be loaded or returned an error (multi-GPU edges only) */
} idsp_status_t;

typedef enum { IDSP_FRAME_MAJOR = 0, IDSP_LANE_MAJOR = 1 } idsp_layout_t;

/* ------------------------------------------------------------------ context */
/* Create a context on `device` with its own non-blocking stream. */
int idsp_b200_init(int device, idsp_ctx **out);
/* Create a context that launches on a caller-owned cudaStream_t (e.g. the
 * current torch stream); the stream must outlive the ctx. */
int idsp_b200_init_on_stream(int device, void *cuda_stream, idsp_ctx **out);
void idsp_b200_free(idsp_ctx *ctx);
int idsp_b200_sync(idsp_ctx *ctx);
const char *idsp_b200_last_error(void);
int idsp_b200_version(void);
/* Number of kernels this ctx has launched so far (bench `gpu_launches`). */
uint64_t idsp_b200_launch_count(const idsp_ctx *ctx);
/* Kernel family of the most recent launch of this ctx, e.g. "tma frame-major wide", "tma lane-major",
 * "generic frame-major", "hbf tiled lane-major", "hbf generic": the fast paths need aligned pointers and
 * whole tiles and fall back silently otherwise, this makes the fallback observable.  Static string. */
const char *idsp_b200_last_kernel(const idsp_ctx *ctx);
/* Page-locked host memory for the `_host` entry points: buffers obtained here (or pinned by
 * the caller with cudaHostRegister / torch pin_memory) are DMA'd directly; pageable memory
 * works too but is staged by the driver. */
int idsp_b200_host_alloc(void **ptr, size_t bytes);
void idsp_b200_host_free(void *ptr);
/* Peer memory for the multi-GPU edges (one process per GPU).  Lanes shard with no collective inside the
 * computation (dsp-process/src/compose.rs:472-475), so the only exchange is handing results to the rank
 * that wants them.  A rank can let its kernels write there directly: the owner allocates a device buffer
 * and exports a 64-byte handle (cudaIpcMemHandle_t), every other process opens the handle and passes
 * the mapped pointer (plus its lane-block offset) as the `y` of any entry point above -- the stores then
 * go over NVLink / NVSwitch from the kernel's epilogue, i.e. the gather is fused into the producer.
 * In the lane-major layout a contiguous lane block of the result is a contiguous range of the buffer.
 * The exporter must keep the allocation alive until every importer has closed it. */
#define IDSP_IPC_HANDLE_BYTES 64
int idsp_b200_malloc(idsp_ctx *ctx, size_t bytes, void **ptr);
int idsp_b200_mfree(idsp_ctx *ctx, void *ptr);
int idsp_b200_ipc_export(idsp_ctx *ctx, const void *ptr, unsigned char handle[IDSP_IPC_HANDLE_BYTES]);
int idsp_b200_ipc_open(idsp_ctx *ctx, const unsigned char handle[IDSP_IPC_HANDLE_BYTES], void **ptr);
int idsp_b200_ipc_close(idsp_ctx *ctx, void *ptr);
/* Copies on the ctx stream for callers that keep samples and state resident on the device between calls
 * (a chained graph, dsp-process/src/compose.rs:13-113, then crosses PCIe once at each end instead of once
 * per stage): kind 0 = host -> device, 1 = device -> host, 2 = device -> device.  Asynchronous like every
 * other call (pinned host memory: idsp_b200_host_alloc); idsp_b200_sync() before reading a download.
 * idsp_b200_memset zero-fills device memory (zero state = the reference's `Default`). */
int idsp_b200_memcpy(idsp_ctx *ctx, void *dst, const void *src, size_t bytes, int kind);
int idsp_b200_memset(idsp_ctx *ctx, void *ptr, int value, size_t bytes);
/* Stream order between two contexts of the same process: everything queued on `signal` so far completes
 * before anything queued on `waiter` after this call starts (an event, no host synchronisation). */
int idsp_b200_stream_wait(idsp_ctx *waiter, idsp_ctx *signal);
/* Kernel selection: 0 = automatic (default), 1 = force the generic LDG kernels,
 * 2 = force the TMA kernels (IDSP_EINVAL if the shape does not qualify),
 * 3 = automatic, with the packed f32x2 variant of the tiled half-band decimator (bit-identical
 * results; the scalar variant is the default because it measured faster). */
int idsp_b200_set_kernel_policy(idsp_ctx *ctx, int policy);

/* ------------------------------------------------------------------ iir::Biquad
 * ba = [b0,b1,b2,a1,a2] raw coefficients exactly as in `Biquad::ba`
 * (src/iir/biquad.rs:96-116).  F = fractional bits of Q<T,A,F> (ignored for
 * floats).  clamp = NULL for `Biquad`, or {u,min,max} for `BiquadClamp`
 * (src/iir/biquad.rs:121-157).
 *
 * DF1: `SplitProcess<T,T,DirectForm1<T>> for Biquad<C>` src/iir/biquad.rs:366-383,
 * clamp :394-404, applied to N lanes by `Lanes<C>` dsp-process/src/compose.rs:468-513.
 * state words: [x[0], x[1], y[0][0], y[0][1]] (DirectForm<T,1,2>, biquad.rs:260-269). */
#define IDSP_DECL_DF1(S, T)                                                          \
    int idsp_biquad_df1_##S(idsp_ctx *ctx, const T ba[5], int F, const T *clamp,     \
                            T *state, const T *x, T *y, size_t frames, size_t lanes, \
                            int layout);                                             \
    int idsp_biquad_df1_##S##_host(idsp_ctx *ctx, const T ba[5], int F,              \
                                   const T *clamp, T *state, const T *x, T *y,       \
                                   size_t frames, size_t lanes, int layout);         \
    /* Cascade<[Biquad<C>;N]> on DirectForm<T,N>: src/iir/biquad.rs:339-364.         \
     * ba = [nsec][5]; state words [x[0],x[1],y[0][0],y[0][1],...,y[N-1][1]];        \
     * 1 <= nsec <= IDSP_MAX_SECTIONS. */                                            \
    int idsp_biquad_cascade_##S(idsp_ctx *ctx, const T *ba, int F, int nsec,         \
                                T *state, const T *x, T *y, size_t frames,           \
                                size_t lanes, int layout);
#define IDSP_MAX_SECTIONS 8
IDSP_DECL_DF1(i8, int8_t)
IDSP_DECL_DF1(i16, int16_t)
IDSP_DECL_DF1(i32, int32_t)
IDSP_DECL_DF1(i64, int64_t)
IDSP_DECL_DF1(f32, float)
IDSP_DECL_DF1(f64, double)

/* DF2T: `SplitProcess<T,T,DirectForm2Transposed<T>> for Biquad<T>` src/iir/biquad.rs:418-440.
 * state words: [x[0], x[1]] (= s0, s1). */
int idsp_biquad_df2t_f32(idsp_ctx *ctx, const float ba[5], const float *clamp, float *state,
                         const float *x, float *y, size_t frames, size_t lanes, int layout);
int idsp_biquad_df2t_f64(idsp_ctx *ctx, const double ba[5], const double *clamp, double *state,
                         const double *x, double *y, size_t frames, size_t lanes, int layout);

/* DirectForm1Wide: src/iir/biquad.rs:445-480; 0 <= F < 32.
 * state words (int32): [x[0], x[1], y[0] lo, y[0] hi, y[1] lo, y[1] hi]. */
int idsp_biquad_df1wide_i32(idsp_ctx *ctx, const int32_t ba[5], int F, const int32_t *clamp,
                            int32_t *state, const int32_t *x, int32_t *y, size_t frames,
                            size_t lanes, int layout);
/* DirectForm1Dither: src/iir/biquad.rs:484-538; 0 <= F < 32.
 * state words (int32): [x[0], x[1], y[0][0], y[0][1], e]. */
int idsp_biquad_df1dither_i32(idsp_ctx *ctx, const int32_t ba[5], int F, const int32_t *clamp,
                              int32_t *state, const int32_t *x, int32_t *y, size_t frames,
                              size_t lanes, int layout);

/* ------------------------------------------------------------------ hbf
 * Built-in taps: HBF_TAPS (src/hbf.rs:308-349), index 0 = lowest rate. */
const float *idsp_hbf_taps(int index, int *M);
size_t idsp_hbf_dec_state_words(int log2_rate);
size_t idsp_hbf_int_state_words(int log2_rate);
#define IDSP_HBF_MAX_M 32

/* /2 decimator `SplitProcess<[T;2],T,HbfDec<[T;N]>> for EvenSymmetric<[C;M]>`
 * src/hbf.rs:155-192.  x: n_out pairs per lane (R = 2), y: n_out per lane.
 * state words: [even history (M-1, oldest first) | odd history (2M-1)]. */
int idsp_hbf_dec_f32(idsp_ctx *ctx, const float *taps, int M, float *state, const float *x,
                     float *y, size_t n_out, size_t lanes, int layout);
/* x2 interpolator `SplitProcess<T,[T;2],HbfInt<[T;N]>>` src/hbf.rs:207-236.
 * state words: [x history (2M-1, oldest first)]. */
int idsp_hbf_int_f32(idsp_ctx *ctx, const float *taps, int M, float *state, const float *x,
                     float *y, size_t n_in, size_t lanes, int layout);
/* HBF_DEC_CASCADE /2^k (src/hbf.rs:385-421; `.inner().1` etc. select the depth):
 * stages TAPS[k-1] -> ... -> TAPS[0].  1 <= log2_rate <= 5.  x: n_out frames of
 * R = 2^k samples per lane, y: n_out per lane.  state = stage states
 * concatenated, highest-rate stage first. */
int idsp_hbf_dec_cascade_f32(idsp_ctx *ctx, int log2_rate, float *state, const float *x,
                             float *y, size_t n_out, size_t lanes, int layout);
int idsp_hbf_dec_cascade_f32_host(idsp_ctx *ctx, int log2_rate, float *state, const float *x,
                                  float *y, size_t n_out, size_t lanes, int layout);
/* HBF_INT_CASCADE x2^k (src/hbf.rs:476-512): stages TAPS[0] -> ... -> TAPS[k-1];
 * state = stage states concatenated, lowest-rate stage first. */
int idsp_hbf_int_cascade_f32(idsp_ctx *ctx, int log2_rate, float *state, const float *x,
                             float *y, size_t n_in, size_t lanes, int layout);
/* The same cascades over CALLER-SUPPLIED half-band tap sets (e.g. HBF_TAPS_98, src/hbf.rs:258-292, or a
 * remez design of the caller): `taps[i]` / `M[i]` = stage i in the order of the reference's tap tuples
 * (index 0 = lowest rate), nstages = log2 of the rate change, 1 <= nstages <= 5, 1 <= M[i] <= IDSP_HBF_MAX_M.
 * Decimator: stages nstages-1 -> 0 (src/hbf.rs:385-421); interpolator: stages 0 -> nstages-1 (:476-512).
 * state = stage states concatenated in processing order, as for the built-in cascades:
 * idsp_hbf_cascade_state_words(decimate, nstages, M) words per lane.  The built-in HBF_TAPS set (compared by
 * value) takes the tiled kernels; any other set runs stage by stage through ctx scratch memory. */
size_t idsp_hbf_cascade_state_words(int decimate, int nstages, const int *M);
int idsp_hbf_dec_cascade_taps_f32(idsp_ctx *ctx, int nstages, const float *const *taps, const int *M,
                                  float *state, const float *x, float *y, size_t n_out, size_t lanes,
                                  int layout);
int idsp_hbf_int_cascade_taps_f32(idsp_ctx *ctx, int nstages, const float *const *taps, const int *M,
                                  float *state, const float *x, float *y, size_t n_in, size_t lanes,
                                  int layout);
/* HBF_TAPS_98 (src/hbf.rs:258-292), index 0 = lowest rate; M = 15, 6, 3, 3, 2. */
const float *idsp_hbf_taps_98(int index, int *M);
/* Single-rate linear-phase FIRs `type_fir!` src/hbf.rs:70-138:
 * odd/sym = (1,1) OddSymmetric, (0,1) EvenSymmetric, (1,0) OddAntiSymmetric,
 * (0,0) EvenAntiSymmetric.  state words: [history (2M-1+odd)]. */
int idsp_fir_f32(idsp_ctx *ctx, const float *taps, int M, int odd, int sym, float *state,
                 const float *x, float *y, size_t frames, size_t lanes, int layout);

/* ------------------------------------------------------------------ cossin / atan2
 * `cossin()` src/cossin.rs:14-67 over an array like idsp.cossin (src/py.rs:11-28):
 * cs[i] = (cos, sin). */
int idsp_cossin_i32(idsp_ctx *ctx, const int32_t *phase, int32_t *cs, size_t n);
int idsp_cossin_i32_host(idsp_ctx *ctx, const int32_t *phase, int32_t *cs, size_t n);
/* `atan2()` src/atan2.rs:66-82 over rows xy[i] = (x, y) like idsp.atan2
 * (src/py.rs:31-46: p[i] = atan2(xy[i][1], xy[i][0])). */
int idsp_atan2_i32(idsp_ctx *ctx, const int32_t *xy, int32_t *p, size_t n);
int idsp_atan2_i32_host(idsp_ctx *ctx, const int32_t *xy, int32_t *p, size_t n);

/* ------------------------------------------------------------------ Lowpass / Lockin
 * `Lowpass<N>` src/lowpass.rs:47-78, order N = 1|2, k = [i32; N] gains.
 * state words (int64): LowpassState<N>.0[0..N]. */
int idsp_lowpass_i32(idsp_ctx *ctx, int order, const int32_t *k, int64_t *state,
                     const int32_t *x, int32_t *y, size_t frames, size_t lanes, int layout);
/* `Lockin<Lowpass<N>>` fed by (sample, phase) src/lockin.rs:30-39 with the phase
 * produced per lane by `Accu` (src/accu.rs:29-38): phase += step before use.
 * accu_state[lanes] (in/out), accu_step[lanes] (device);
 * lp_state words (int64): [I state (N) | Q state (N)];
 * iq = Complex<i32> per sample: flat index of x, times 2, + {0: re, 1: im}. */
int idsp_lockin_i32(idsp_ctx *ctx, int order, const int32_t *k, int32_t *accu_state,
                    const int32_t *accu_step, int64_t *lp_state, const int32_t *x, int32_t *iq,
                    size_t frames, size_t lanes, int layout);
int idsp_lockin_i32_host(idsp_ctx *ctx, int order, const int32_t *k, int32_t *accu_state,
                         const int32_t *accu_step, int64_t *lp_state, const int32_t *x,
                         int32_t *iq, size_t frames, size_t lanes, int layout);
/* `Lockin<Lowpass<N>>` on (sample, phase) tuples, `SplitProcess<(i32, Wrapping<i32>), Complex<i32>, [S; 2]>`
 * src/lockin.rs:30-39: the phase comes from the caller (e.g. the PLL output, src/pll.rs:89-108) instead
 * of a per-lane `Accu`.  xp = frames * lanes (x, phase) i32 pairs in the layout of x (pair innermost),
 * 8-byte aligned; lp_state as above. */
int idsp_lockin_phase_i32(idsp_ctx *ctx, int order, const int32_t *k, int64_t *lp_state, const int32_t *xp,
                          int32_t *iq, size_t frames, size_t lanes, int layout);
/* `Lockin<Lowpass<N>>` on (sample, LO) tuples, `SplitProcess<(X, Complex<U>), Complex<X>, [S; 2]>`
 * src/lockin.rs:17-28 with X = i32, U = Q32<32> (`i32 * Q32<32>` = (x * lo) >> 32,
 * dsp-fixedpoint/src/lib.rs:449-456): xlo = frames * lanes (x, lo.re, lo.im) i32 triples. */
int idsp_lockin_lo_i32(idsp_ctx *ctx, int order, const int32_t *k, int64_t *lp_state, const int32_t *xlo,
                       int32_t *iq, size_t frames, size_t lanes, int layout);

/* ------------------------------------------------------------------ fused chain
 * HbfDec(/2^k) -> HbfInt(x2^k) -> Biquad<f32> DF1 in one pass (BASELINE config 5;
 * composition of hbf.rs:385-421, :476-512 and biquad.rs:366-383).
 * state = [dec cascade state | int cascade state | DF1 state (4)].
 * x, y: n_low frames of 2^k samples per lane. */
size_t idsp_chain_state_words(int log2_rate);
int idsp_chain_f32(idsp_ctx *ctx, int log2_rate, const float ba[5], float *state,
                   const float *x, float *y, size_t n_low, size_t lanes, int layout);
/* host buffers: one PCIe round trip carries the three operators */
int idsp_chain_f32_host(idsp_ctx *ctx, int log2_rate, const float ba[5], float *state,
                        const float *x, float *y, size_t n_low, size_t lanes, int layout);

/* ------------------------------------------------------------------ cic::Cic
 * `Cic<T, N, M>` src/cic.rs:13-200 (order N = 1..6, comb delay M = 1..3, rate = fast/slow - 1) under
 * the chunk adapters of dsp-process/src/adapters.rs: `Decimator` (:154-222, `[T; rate+1] -> T`, the
 * value of the frame's single tick) and `Interpolator` (:27-35, `T -> [T; rate+1]`).  Integer
 * arithmetic wraps (wrapping_add / wrapping_sub in the decimator, release-mode `+=` / `-` in the
 * interpolator).  State: SoA words of T, `state[w * lanes + lane]`, w = [index, zoh,
 * combs[n][m] at 2 + n*M + m, integrators[n] at 2 + N*M + n]; all zero = `Cic::new(rate)`.
 * Whole frames only, so `index` is 0 between calls (the adapters' one-tick-per-chunk contract).
 * dec: x = frames * lanes * (rate+1) samples, y = frames * lanes; int: the reverse. */
size_t idsp_cic_state_words(int N, int M);
int idsp_cic_dec_i32(idsp_ctx *ctx, int N, int M, uint32_t rate, int32_t *state, const int32_t *x,
                     int32_t *y, size_t frames, size_t lanes, int layout);
int idsp_cic_dec_i64(idsp_ctx *ctx, int N, int M, uint32_t rate, int64_t *state, const int64_t *x,
                     int64_t *y, size_t frames, size_t lanes, int layout);
int idsp_cic_int_i32(idsp_ctx *ctx, int N, int M, uint32_t rate, int32_t *state, const int32_t *x,
                     int32_t *y, size_t frames, size_t lanes, int layout);
int idsp_cic_int_i64(idsp_ctx *ctx, int N, int M, uint32_t rate, int64_t *state, const int64_t *x,
                     int64_t *y, size_t frames, size_t lanes, int layout);

/* ------------------------------------------------------------------ pll::PLL
 * `SplitProcess<W<i32>, W<i32>, PLLState> for PLL` src/pll.rs:88-108 (phase clamp: `ClampWrap`,
 * src/unwrap.rs:166-194).  ba = the three raw `Q32<32>` lead-lag coefficients (`PLL::ba`, e.g. from
 * `PLL::from_bandwidth`, src/pll.rs:41-57); x = input phase, y = output phase estimate (`state.y`).
 * State: SoA i32 words [clamp.x0, clamp.clamp (-1|0|1), z0, y0, f0 lo, f0 hi, f lo, f hi, y];
 * `PLLState::frequency()` is word 7.  All arithmetic wraps, bit-exact. */
int idsp_pll_i32(idsp_ctx *ctx, const int32_t *ba, int32_t *state, const int32_t *x, int32_t *y,
                 size_t frames, size_t lanes, int layout);

/* ------------------------------------------------------------------ FM discriminator
 * The fixed-point core of examples/fm_disc.rs:26-48 fused into one pass per lane:
 * `z = x * prev.into_bits().conj()` (`Complex<Q32<32>> * Complex<i32>`, src/complex.rs:117-134),
 * `d = z.arg() - carrier` (src/complex.rs:254-256, wrapping), `y = Biquad<Q32<F>>` DF1 of d
 * (src/iir/biquad.rs:366-383, ba = 5 raw coefficients).  x = frames * lanes (re, im) i32 pairs,
 * y = frames * lanes.  State: SoA i32 words [has_prev, prev.re, prev.im, x1, x2, y1, y2];
 * all zero = `Split::new(FmDiscriminator{..}, None)` * `DirectForm1::default()`. */
int idsp_fm_disc_i32(idsp_ctx *ctx, int32_t carrier, const int32_t *ba, int F, int32_t *state,
                     const int32_t *x, int32_t *y, size_t frames, size_t lanes, int layout);

/* ------------------------------------------------------------------ multi-GPU edges
 * One process per GPU.  Lanes shard as contiguous blocks (rank r owns idsp_b200_lane_block(...)) with no
 * collective inside the computation (dsp-process/src/compose.rs:472-475: lanes never interact); the two
 * edges -- handing lane blocks of a root-resident buffer to the ranks, collecting the results -- are grouped
 * NCCL point-to-point transfers over NVLink / NVSwitch on the ctx stream (asynchronous like every other
 * call).  They sit under `Split<Lanes<C>, [S; N]>` (dsp-process/src/split.rs:272-277): scatter x, run any
 * entry point above on the local block, gather y.  NCCL is loaded at run time (libnccl.so.2); a single
 * rank communicator (nranks == 1) never touches it.  The fused alternative to the gather is the peer
 * memory above (kernels store their results straight into the root's buffer).
 *  - idsp_b200_comm_unique_id: rank 0 creates the 128-byte id and hands it to the other processes by any
 *    out-of-band means (file, socket, MPI, torch.distributed object broadcast);
 *  - `elem_bytes` = bytes per frame and lane (4 for i32 / f32 samples, 8 for Complex<i32>, 64 for [f32; 16]);
 *  - `full` is only read / written on `root` (may be NULL elsewhere), `part` holds this rank's block in the
 *    same layout with its own lane count (hi - lo). */
#define IDSP_COMM_ID_BYTES 128
typedef struct idsp_comm idsp_comm;
int idsp_b200_comm_unique_id(unsigned char id[IDSP_COMM_ID_BYTES]);
int idsp_b200_comm_init(idsp_ctx *ctx, int nranks, int rank, const unsigned char id[IDSP_COMM_ID_BYTES],
                        idsp_comm **out);
int idsp_b200_comm_free(idsp_comm *comm);
int idsp_b200_comm_rank(const idsp_comm *comm);
int idsp_b200_comm_size(const idsp_comm *comm);
int idsp_b200_nccl_version(void); /* 0 if NCCL cannot be loaded */
/* lanes [lo, hi) of `rank`: whole units of `align` lanes (0 = 32, a warp), covering [0, lanes) exactly */
int idsp_b200_lane_block(size_t lanes, int nranks, int rank, size_t align, size_t *lo, size_t *hi);
int idsp_scatter_lanes(idsp_comm *comm, const void *full, void *part, size_t frames, size_t lanes,
                       size_t elem_bytes, int layout, int root);
int idsp_gather_lanes(idsp_comm *comm, const void *part, void *full, size_t frames, size_t lanes,
                      size_t elem_bytes, int layout, int root);
/* replicated small data (coefficients): root's bytes to every rank, in place, device memory */
int idsp_broadcast(idsp_comm *comm, void *buf, size_t bytes, int root);
/* Building blocks for pipelined edges (a lane block handed over in sub-blocks while the previous one is being
 * filtered): plain point-to-point transfers on the communicator's ctx stream, grouped into one NCCL launch
 * between idsp_comm_group_begin / _end.  Give the communicator its own ctx (its own stream) and order the
 * compute ctx behind it with idsp_b200_stream_wait(). */
int idsp_comm_group_begin(idsp_comm *comm);
int idsp_comm_group_end(idsp_comm *comm);
int idsp_comm_send(idsp_comm *comm, const void *buf, size_t bytes, int peer);
int idsp_comm_recv(idsp_comm *comm, void *buf, size_t bytes, int peer);

/* ------------------------------------------------------------------ coefficient builders
 * Host-side, no device work: the reference's `iir::coefficients::Filter` (src/iir/coefficients.rs:111-527),
 * `Biquad::from([[T; 3]; 2])` / `from([T; 5])` / `from_zpk` (src/iir/biquad.rs:545-619) and
 * `pid::Builder::build` (src/iir/pid.rs:236-317), each in the reference's two float widths (the f32 impl
 * rounds every intermediate to f32, like `Filter<f32>` -> `Biquad<Q32<30>>` in examples/fm_disc.rs).
 * Quantisation to Q<T, A, F>: (v * 2^F).round() (half away from zero) then Rust `as` (saturating, NaN -> 0),
 * dsp-fixedpoint/src/num_traits_impl.rs:32-45.  Validation errors (IDSP_EINVAL) carry the reference's
 * `iir::Error` variant and field in idsp_b200_last_error(), e.g. "OutOfRange(frequency)". */
typedef enum { IDSP_I8 = 0, IDSP_I16 = 1, IDSP_I32 = 2, IDSP_I64 = 3, IDSP_F32 = 4, IDSP_F64 = 5 } idsp_kind_t;
typedef enum { /* coefficients::Type, src/iir/coefficients.rs:44-66 */
    IDSP_LOWPASS = 0, IDSP_HIGHPASS = 1, IDSP_BANDPASS = 2, IDSP_ALLPASS = 3, IDSP_NOTCH = 4,
    IDSP_PEAKING = 5, IDSP_LOWSHELF = 6, IDSP_HIGHSHELF = 7, IDSP_IHO = 8
} idsp_filter_type_t;
typedef enum { IDSP_SHAPE_Q = 0, IDSP_SHAPE_BANDWIDTH = 1, IDSP_SHAPE_SLOPE = 2 } idsp_shape_t;
/* `Filter<T>` (coefficients.rs:28-41); Default = {0, 1, 1, Q(1/sqrt 2)} */
typedef struct { double frequency, gain, shelf; int shape_kind; double shape; } idsp_filter_f64;
typedef struct { float frequency, gain, shelf; int shape_kind; float shape; } idsp_filter_f32;
void idsp_filter_default_f64(idsp_filter_f64 *f);
void idsp_filter_default_f32(idsp_filter_f32 *f);
int idsp_filter_validate_f64(const idsp_filter_f64 *f); /* coefficients.rs:241-265 */
int idsp_filter_validate_f32(const idsp_filter_f32 *f);
/* `Filter::build(typ)` (coefficients.rs:466-479): ba6 = [b0, b1, b2, a0, a1, a2], literature sign of a1 / a2 */
int idsp_filter_build_f64(const idsp_filter_f64 *f, int type, double ba6[6]);
int idsp_filter_build_f32(const idsp_filter_f32 *f, int type, float ba6[6]);
/* `Biquad<C>::from([[T; 3]; 2])` (biquad.rs:545-566): normalise by a0, flip the sign of a1 / a2, convert to
 * the coefficient type `kind` (Q format with F fractional bits, or f32 / f64): out = 5 values of that type. */
int idsp_biquad_from_ba6_f64(const double ba6[6], int kind, int F, void *ba5_out);
int idsp_biquad_from_ba6_f32(const float ba6[6], int kind, int F, void *ba5_out);
/* `Biquad<C>::from([T; 5])` (biquad.rs:568-576): conversion only */
int idsp_biquad_from_ba5_f64(const double ba5[5], int kind, int F, void *ba5_out);
int idsp_biquad_from_ba5_f32(const float ba5[5], int kind, int F, void *ba5_out);
/* `Biquad::from_zpk` (biquad.rs:594-619): pairs are (x, y); complex != 0 means the conjugate pair x +- jy */
int idsp_biquad_from_zpk_f64(const double zeros[2], int zeros_complex, const double poles[2], int poles_complex,
                             double gain, int kind, int F, void *ba5_out);
/* `Filter::build_biquad` / `try_build_biquad` (coefficients.rs:481-497) in one call */
int idsp_filter_build_biquad_f64(const idsp_filter_f64 *f, int type, int kind, int F, void *ba5_out);
int idsp_filter_build_biquad_f32(const idsp_filter_f32 *f, int type, int kind, int F, void *ba5_out);
/* `pid::Builder<T>` (src/iir/pid.rs:38-47): order = pid::Order (P = 2, I = 1, I2 = 0), gain / limit indexed
 * by pid::Action (I2, I, P, D, D2); Default = {I, 0.., +inf..}.  build(period): pid.rs:236-303, the GAINS are
 * converted to the coefficient type and accumulated there (wrapping for Q formats). */
typedef struct { int order; double gain[5], limit[5]; } idsp_pid_f64;
typedef struct { int order; float gain[5], limit[5]; } idsp_pid_f32;
void idsp_pid_default_f64(idsp_pid_f64 *b);
void idsp_pid_default_f32(idsp_pid_f32 *b);
int idsp_pid_validate_f64(const idsp_pid_f64 *b, double period); /* pid.rs:193-222 */
int idsp_pid_validate_f32(const idsp_pid_f32 *b, float period);
int idsp_pid_build_f64(const idsp_pid_f64 *b, double period, int kind, int F, void *ba5_out);
int idsp_pid_build_f32(const idsp_pid_f32 *b, float period, int kind, int F, void *ba5_out);

#ifdef __cplusplus
}
#endif
#endif /* IDSP_B200_H */
