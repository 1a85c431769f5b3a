/*
 * idsp_oracle.h -- CPU restatement ("oracle") of the quartiq/idsp filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under idsp_b200/ (the product) may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * The reference is pure Rust and cannot be built in this image (no rustc/cargo),
 * so this is a plain-C restatement of its arithmetic.  Each function cites the
 * reference file:line it follows (paths relative to the reference tree).
 * Parity pinning: tests/test_oracle_kat.py checks every known-answer test the
 * reference carries for this path.  Lowpass/Lockin and
 * DirectForm1Wide have no value-level test in the reference: for those the
 * header says "parity unpinned" (cross-checked against an independent Python
 * big-integer model in tests/pymodel.py instead).
 *
 * Conventions
 *  - all integer accumulations wrap (Rust release semantics), shifts on signed
 *    values are arithmetic, float ops are individually rounded (compile with
 *    -ffp-contract=off, no fast-math), denormals kept.
 *  - multi-lane state is SoA: state[word * lanes + lane]; word order = the
 *    reference struct's field order (see each function).
 *  - layout: 0 = frame-major flat[t*lanes + l] (dsp-process/src/view.rs:106-131),
 *            1 = lane-major  flat[l*frames + t] (dsp-process/src/view.rs:176-225).
 */
#ifndef IDSP_ORACLE_H
#define IDSP_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_FRAME_MAJOR 0
#define ORC_LANE_MAJOR 1

/* ---- tables (build.rs:9-69), generated at first use with libm ---- */
const uint32_t *orc_cossin_table(void);          /* [128] */
const uint32_t *orc_atan2_divi_base(void);       /* [16]  */
const int32_t *orc_atan2_divi_slope(void);       /* [16]  */

/* ---- memoryless ---- */
void orc_cossin(int32_t phase, int32_t *cos_out, int32_t *sin_out); /* src/cossin.rs:14-67 */
int32_t orc_atan2(int32_t y, int32_t x);                            /* src/atan2.rs:66-82 */
void orc_cossin_n(const int32_t *phase, int32_t *cs /*[n][2]*/, size_t n);  /* src/py.rs:11-28 */
void orc_atan2_n(const int32_t *xy /*[n][2]=x,y*/, int32_t *p, size_t n);   /* src/py.rs:31-46 */

/* ---- coefficient quantisation (src/iir/biquad.rs:545-576, num_traits_impl.rs:32-45) ---- */
/* ba6 = [b0,b1,b2,a0,a1,a2] (literature signs) -> normalised, sign-flipped
 * [b0,b1,b2,a1,a2]/a0 in f64 */
void orc_ba_normalize_f64(const double ba6[6], double out5[5]);
int8_t orc_quant_i8(double v, int F);
int16_t orc_quant_i16(double v, int F);
int32_t orc_quant_i32(double v, int F);
int64_t orc_quant_i64(double v, int F);
int32_t orc_round_sat_i32(double v); /* f64::round() as i32 (src/py.rs:99-101) */

/* ---- single-lane biquads: st words documented per function ---- */
/* DF1 (src/iir/biquad.rs:366-383), clamp (:394-404). st = [x0,x1,y0,y1].
 * clamp == NULL -> plain Biquad; else {u,min,max}. */
#define ORC_DECL_DF1(S, T)                                                        \
    void orc_biquad_df1_##S(const T ba[5], int F, const T *clamp, T st[4],        \
                            const T *x, T *y, size_t n);                          \
    void orc_biquad_df1_##S##_lanes(const T ba[5], int F, const T *clamp,         \
                                    T *st /*[4][lanes]*/, const T *x, T *y,       \
                                    size_t frames, size_t lanes, int layout,      \
                                    int nthreads);                                \
    /* Cascade<[Biquad;N]> on DirectForm<T,N> (biquad.rs:339-364):                \
     * ba = [nsec][5], st = [x0,x1,y[0][0],y[0][1],...,y[N-1][1]] (2+2N words) */ \
    void orc_biquad_cascade_##S(const T *ba, int F, int nsec, T *st, const T *x,  \
                                T *y, size_t n);                                  \
    void orc_biquad_cascade_##S##_lanes(const T *ba, int F, int nsec, T *st,      \
                                        const T *x, T *y, size_t frames,          \
                                        size_t lanes, int layout, int nthreads);
ORC_DECL_DF1(i8, int8_t)
ORC_DECL_DF1(i16, int16_t)
ORC_DECL_DF1(i32, int32_t)
ORC_DECL_DF1(i64, int64_t)
ORC_DECL_DF1(f32, float)
ORC_DECL_DF1(f64, double)

/* DF2T (src/iir/biquad.rs:418-440). st = [s0,s1]. */
#define ORC_DECL_DF2T(S, T)                                                       \
    void orc_biquad_df2t_##S(const T ba[5], const T *clamp, T st[2], const T *x,  \
                             T *y, size_t n);                                     \
    void orc_biquad_df2t_##S##_lanes(const T ba[5], const T *clamp, T *st,        \
                                     const T *x, T *y, size_t frames,             \
                                     size_t lanes, int layout, int nthreads);
ORC_DECL_DF2T(f32, float)
ORC_DECL_DF2T(f64, double)

/* DirectForm1Wide (src/iir/biquad.rs:445-480) -- parity unpinned in the reference.
 * st (int32 words) = [x0,x1,y0_lo,y0_hi,y1_lo,y1_hi] */
void orc_biquad_df1wide_i32(const int32_t ba[5], int F, const int32_t *clamp,
                            int32_t st[6], const int32_t *x, int32_t *y, size_t n);
void orc_biquad_df1wide_i32_lanes(const int32_t ba[5], int F, const int32_t *clamp,
                                  int32_t *st, const int32_t *x, int32_t *y,
                                  size_t frames, size_t lanes, int layout, int nthreads);
/* DirectForm1Dither (src/iir/biquad.rs:484-538). st = [x0,x1,y0,y1,e] */
void orc_biquad_df1dither_i32(const int32_t ba[5], int F, const int32_t *clamp,
                              int32_t st[5], const int32_t *x, int32_t *y, size_t n);
void orc_biquad_df1dither_i32_lanes(const int32_t ba[5], int F, const int32_t *clamp,
                                    int32_t *st, const int32_t *x, int32_t *y,
                                    size_t frames, size_t lanes, int layout, int nthreads);

/* ---- python-FFI shaped entry points (src/py.rs:50-108) ---- */
void orc_sos(const double *sos /*[nsec][6]*/, int nsec, int32_t *xy, size_t n);
void orc_sos_clamp_wide(const double *sos /*[nsec][9]*/, int nsec, int32_t *xy, size_t n);

/* ---- half band filters (src/hbf.rs) ---- */
#define ORC_HBF_MAX_M 64
extern const float ORC_HBF_TAPS0[23], ORC_HBF_TAPS1[10], ORC_HBF_TAPS2[5],
    ORC_HBF_TAPS3[4], ORC_HBF_TAPS4[3];                      /* hbf.rs:308-349 */
const float *orc_hbf_taps(int idx, int *M);
/* words of state per lane of the /2^k and x2^k cascades */
size_t orc_hbf_dec_state_words(int log2_rate);
size_t orc_hbf_int_state_words(int log2_rate);
size_t orc_hbf_dec_response_length(int depth); /* hbf.rs:424-448 */
size_t orc_hbf_int_response_length(int depth); /* hbf.rs:515-539 */

/* Single /2 stage (hbf.rs:155-192): x = n pairs [even,odd], y = n outputs.
 * st = [even_hist (M-1, oldest first) | odd_hist (2M-1, oldest first)] */
void orc_hbf_dec_f32(const float *taps, int M, float *st, const float *x, float *y, size_t n);
/* Single x2 stage (hbf.rs:207-236): x = n inputs, y = n pairs. st = [x_hist (2M-1)] */
void orc_hbf_int_f32(const float *taps, int M, float *st, const float *x, float *y, size_t n);
/* Cascades (hbf.rs:385-421, 476-512). st = concatenation of the stage states,
 * highest-rate stage first for dec, lowest-rate stage first for int.
 * dec: x = n_out*2^k inputs, y = n_out.  int: x = n_in, y = n_in*2^k. */
void orc_hbf_dec_cascade_f32(int log2_rate, float *st, const float *x, float *y, size_t n_out);
void orc_hbf_int_cascade_f32(int log2_rate, float *st, const float *x, float *y, size_t n_in);
/* Lanes: frame-major x[t][lane][R] / y[t][lane]; lane-major x[lane][t*R..] */
void orc_hbf_dec_cascade_f32_lanes(int log2_rate, float *st, const float *x, float *y,
                                   size_t n_out, size_t lanes, int layout, int nthreads);
void orc_hbf_int_cascade_f32_lanes(int log2_rate, float *st, const float *x, float *y,
                                   size_t n_in, size_t lanes, int layout, int nthreads);
/* generic single-rate symmetric FIRs (hbf.rs:70-138): odd/sym select the type.
 * st = [hist (2M-1+odd)] */
void orc_fir_f32(const float *taps, int M, int odd, int sym, float *st, const float *x,
                 float *y, size_t n);

/* ---- Lowpass / Lockin (src/lowpass.rs:47-78, src/lockin.rs:17-39) -- parity unpinned ---- */
/* order 1|2; st = int64[order] */
void orc_lowpass_i32(int order, const int32_t *k, int64_t *st, const int32_t *x, int32_t *y,
                     size_t n);
void orc_lowpass_i32_lanes(int order, const int32_t *k, int64_t *st /*[order][lanes]*/,
                           const int32_t *x, int32_t *y, size_t frames, size_t lanes,
                           int layout, int nthreads);
/* Accu (src/accu.rs:29-38) + Lockin<Lowpass<order>> fed by phase.
 * per lane: accu state/step (i32, wrapping), lp state int64 [2][order] (I then Q).
 * x[t][lane] (frame-major) or x[lane][t]; iq same indexing with [2] innermost. */
void orc_lockin_i32_lanes(int order, const int32_t *k, int32_t *accu_state,
                          const int32_t *accu_step, int64_t *lp_st /*[2*order][lanes]*/,
                          const int32_t *x, int32_t *iq, size_t frames, size_t lanes,
                          int layout, int nthreads);
/* Lockin on (sample, phase) tuples (lockin.rs:30-39) and on (sample, LO) tuples (lockin.rs:17-28) */
void orc_lockin_phase_i32_lanes(int order, const int32_t *k, int64_t *lp_st, const int32_t *xp, int32_t *iq,
                                size_t frames, size_t lanes, int layout, int nthreads);
void orc_lockin_lo_i32_lanes(int order, const int32_t *k, int64_t *lp_st, const int32_t *xlo, int32_t *iq,
                             size_t frames, size_t lanes, int layout, int nthreads);

/* f32 chain of config 5: HbfDec(/2^k) -> HbfInt(x2^k) -> Biquad DF1 f32.
 * st = [dec state | int state | df1 state(4)] words per lane (SoA over lanes). */
void orc_chain_f32_lanes(int log2_rate, const float ba[5], float *st, const float *x, float *y,
                         size_t n_low /*low-rate frames*/, size_t lanes, int layout,
                         int nthreads);

int orc_max_threads(void);

/* ---- Cic<T,N,M> (src/cic.rs:13-200) with the Decimator / Interpolator adapters
 * (dsp-process/src/adapters.rs:27-35, :154-222); state words [index, zoh, combs[N][M], integrators[N]] */
size_t orc_cic_state_words(int N, int M);
void orc_cic_dec_i32_lanes(int N, int M, uint32_t rate, int32_t *st, const int32_t *x, int32_t *y,
                           size_t frames, size_t lanes, int layout, int nthreads);
void orc_cic_dec_i64_lanes(int N, int M, uint32_t rate, int64_t *st, const int64_t *x, int64_t *y,
                           size_t frames, size_t lanes, int layout, int nthreads);
void orc_cic_int_i32_lanes(int N, int M, uint32_t rate, int32_t *st, const int32_t *x, int32_t *y,
                           size_t frames, size_t lanes, int layout, int nthreads);
void orc_cic_int_i64_lanes(int N, int M, uint32_t rate, int64_t *st, const int64_t *x, int64_t *y,
                           size_t frames, size_t lanes, int layout, int nthreads);
int64_t orc_cic_gain(int N, int M, uint32_t rate);
uint32_t orc_cic_gain_log2(int N, int M, uint32_t rate);
size_t orc_cic_response_length(int N, uint32_t rate);

/* ---- PLL (src/pll.rs:33-108, ClampWrap src/unwrap.rs:166-194); state words (i32):
 * [x0, clamp, z0, y0, f0 lo, f0 hi, f lo, f hi, y]; frequency() = f hi */
void orc_pll_from_bandwidth(float bw, float split, int32_t ba[3]);
void orc_pll_i32_lanes(const int32_t ba[3], int32_t *st /*[9][lanes]*/, const int32_t *x, int32_t *y,
                       size_t frames, size_t lanes, int layout, int nthreads);

/* ---- FM discriminator graph (examples/fm_disc.rs:26-48): x = (re, im) i32 pairs; state words (i32)
 * [has_prev, prev.re, prev.im, x1, x2, y1, y2] */
void orc_fm_disc_i32_lanes(int32_t carrier, const int32_t ba[5], int F, int32_t *st, const int32_t *x,
                           int32_t *y, size_t frames, size_t lanes, int layout, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
