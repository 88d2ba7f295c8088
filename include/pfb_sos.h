/*
 * pfb_sos.h -- C ABI of libpfb_sos.so: B200 (sm_100a) kernels for pyfilterbank's cascaded
 * second-order-section (biquad) IIR filtering hot path.
 *
 * Plain C, plain pointers and sizes; no torch / C++ types cross this boundary.
 *
 * Layouts (identical to the reference, pyfilterbank/sosfilt.c + sosfiltering.py):
 *   sos     : rows of 6 doubles [b0 b1 b2 a0 a1 a2] per section; a0 is read and ignored
 *             (sosfilt.c:13,39,67).  A bank is [B][K][6] (shared by all channels) or
 *             [C][B][K][6] (per channel, the order sosfilter_double_mimo consumes, sosfilt.c:63-68).
 *   states  : [C][B][K][2] = (w1, w2) Direct-Form-II delay values per section (sosfilt.c:61-62,79-80).
 *   signals : input x [C][ldx] (channel rows, time contiguous); output y [C][B][ldy]
 *             (== the reference's (N, B, C) Fortran-ordered result, sosfiltering.py:248).
 *
 * Every function that can fail returns 0 on success or a negative pfb_status; the text of the
 * last failure on the calling thread is available from pfb_last_error().  The three reference
 * drop-ins return void like the originals: on failure they print the error to stderr, leave the
 * buffers untouched and record the status for pfb_last_status().
 */
#ifndef PFB_SOS_H
#define PFB_SOS_H

#include <stdint.h>

#if defined(__GNUC__)
#define PFB_API __attribute__((visibility("default")))
#else
#define PFB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pfb_status {
    PFB_OK = 0,
    PFB_ERR_INVALID = -1,   /* bad argument */
    PFB_ERR_CUDA = -2,      /* CUDA runtime error, see pfb_last_error() */
    PFB_ERR_NO_DEVICE = -3, /* no CUDA device: there is no CPU fallback */
    PFB_ERR_ALLOC = -4
} pfb_status;

/* element types of signal buffers */
#define PFB_F32 0
#define PFB_F64 1

/* flags for pfb_bank_filter() */
#define PFB_MODE_AUTO 0x0   /* pick below from (cascades, samples) */
#define PFB_MODE_SEQ 0x1    /* one thread per (channel, band) cascade, sequential in time */
#define PFB_MODE_SCAN 0x2   /* time-chunked: zero-state pre-pass, state-transition carry scan, output pass */
#define PFB_MODE_MASK 0xF
#define PFB_EXACT 0x10      /* no FMA contraction, the reference's operation order: with PFB_MODE_SEQ the
                               result is bit-identical to sosfilt.c */
#define PFB_ARITH_F64 0x20  /* float32 signals, float64 coefficients/states/arithmetic */
#define PFB_DECIMATE2 0x40  /* store only every second output sample: y rows hold (n - phase + 1) / 2 samples,
                               output i = filtered sample 2 i + phase (the anti-alias stage of the decimating
                               band path; banks with K in {1,2,3,4,6,8} only) */
#define PFB_DEC_PHASE1 0x80 /* phase = 1 (default 0): the block starts on an odd sample of the decimator */

/* ------------------------------------------------------------------------------------------------
 * 1. Reference drop-ins: the three symbols pyfilterbank/build.py:5-7 declares and
 *    sosfiltering.py:92,149,228 call through cffi.  HOST pointers, in place, same argument meaning.
 *    Arithmetic: sequential per cascade, separately rounded mul/add in the reference's order
 *    (PFB_MODE_SEQ | PFB_EXACT), so results are bit-identical to the CPU reference.
 * ---------------------------------------------------------------------------------------------- */
PFB_API void sosfilter(float *signal, int nsamp, float *sos, int ksos, float *states);                 /* sosfilt.c:5  */
PFB_API void sosfilter_double(double *signal, int nsamp, double *sos, int ksos, double *states);       /* sosfilt.c:31 */
PFB_API void sosfilter_double_mimo(double *signal, int nframes, int nchan, double *sos, int ksos,      /* sosfilt.c:57 */
                           int kbands, double *states);

PFB_API int pfb_last_status(void);
PFB_API const char *pfb_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * 2. Bank objects: device-resident coefficient tables for one filter bank (what
 *    FractionalOctaveFilterbank.sosmat holds, octbank.py:374-386).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pfb_bank pfb_bank;

/* sos: HOST [rows][K][6] doubles, rows == B (shared) or C*B (per_channel != 0).
 * The bank is created on the current CUDA device. */
PFB_API int pfb_bank_create(pfb_bank **bank, const double *sos, int C_sos, int B, int K, int per_channel);
PFB_API void pfb_bank_destroy(pfb_bank *bank);

/* Filter C channels through the bank.  DEVICE pointers; asynchronous on `stream` (a cudaStream_t).
 *   dtype   : PFB_F32 or PFB_F64, element type of x, y and states (states are float64 when
 *             PFB_ARITH_F64 is set for float32 signals)
 *   x       : [C][ldx], y : [C][B][ldy]; x may alias y only when B == 1 and ldx == ldy (in place)
 *   states  : [C][B][K][2], read for the initial state and overwritten with the final one
 *   n       : samples per channel (64-bit) */
PFB_API int pfb_bank_filter(pfb_bank *bank, int dtype, const void *x, int64_t ldx, void *y, int64_t ldy,
                    void *states, int64_t n, int C, int flags, void *stream);

/* Band-level epilogue: the band signals are reduced to energies in the kernel instead of being written --
 * levels[(c * B + b) * ld_levels + k] = sum of y[c][b][t]^2 over block k = [k * block_len, (k + 1) * block_len)
 * (what the reference's users compute from filter_mimo_c's output, `L = 10*log10(sum(y**2))`, octbank.py:262-263,
 * :536-537).  No band signal touches HBM.  DEVICE pointers, asynchronous on `stream`; shared banks only;
 * n must be a multiple of block_len and block_len a multiple of 32 (float64) / 64 (float32) samples.
 * States are carried like in pfb_bank_filter. */
PFB_API int pfb_bank_levels(pfb_bank *bank, int dtype, const void *x, int64_t ldx, double *levels, int64_t ld_levels,
                            void *states, int64_t n, int C, int64_t block_len, int flags, void *stream);

/* pfb_bank_levels on HOST buffers: the input streams through the device in channel groups, only the energies come back
 * (levels [C*B][ld_levels] on the host).  Synchronous. */
PFB_API int pfb_bank_levels_host(pfb_bank *bank, int dtype, const void *x, int64_t ldx, double *levels, int64_t ld_levels,
                                 void *states, int64_t n, int C, int64_t block_len, int flags);

/* Same on HOST buffers (pageable or pinned): streams the signal through the device in time blocks
 * with carried states, overlapping H2D, kernels and D2H.  Synchronous. */
PFB_API int pfb_bank_filter_host(pfb_bank *bank, int dtype, const void *x, int64_t ldx, void *y, int64_t ldy,
                         void *states, int64_t n, int C, int flags);

/* Time-chunked mode hands every chunk the zero-state response to the last W samples of its
 * predecessor (plus the exactly propagated older state).  W is chosen per band from the band's impulse
 * response so that the dropped tail stays below `tol` relative L2 (default 1e-11, two orders under the
 * float64 parity budget; float32 arithmetic uses max(tol, 1e-6), two orders under its 1e-4 budget; env
 * PFB_WARM_TOL overrides).  tol = 0 forces W = chunk length (no truncation). */
PFB_API int pfb_bank_set_warm_tol(pfb_bank *bank, double tol);

/* Introspection used by the bench harness: what the last pfb_bank_filter on this bank launched. */
typedef struct pfb_launch_info {
    int mode;             /* PFB_MODE_SEQ or PFB_MODE_SCAN */
    int kernels;          /* kernel launches issued */
    int warps_per_cascade;
    int64_t chunk;        /* samples per time chunk (scan mode), n in seq mode */
    int64_t warm_total;   /* sum over cascades of pre-pass samples per chunk (scan mode) */
    int aligned;          /* 0 / 1: cp.async kernels without / with 16-byte vectors; 2: TMA kernels; 3: TMA kernels
                             with the 512-byte-fragment output pass */
} pfb_launch_info;
PFB_API int pfb_bank_last_launch(const pfb_bank *bank, pfb_launch_info *info);

/* Per-kernel device times for the bench harness.  timing_begin arms CUDA events for the next
 * `max_calls` pfb_bank_filter calls on this bank (recorded on the stream each call is launched on);
 * timing_end waits for them and returns the mean milliseconds per call of
 *   ms[0] pre-pass kernel, ms[1] carry kernel, ms[2] output kernel, ms[3] whole call
 * (paths that launch a single kernel book everything on ms[2]) and how many calls were recorded. */
PFB_API int pfb_bank_timing_begin(pfb_bank *bank, int max_calls);
PFB_API int pfb_bank_timing_end(pfb_bank *bank, double *ms, int *calls);

/* ------------------------------------------------------------------------------------------------
 * 3. The two scipy.signal.lfilter recurrences on the path.  DEVICE pointers, asynchronous on `stream`.
 * ---------------------------------------------------------------------------------------------- */

/* GammatoneFilterbank.analyze / fosfilter (gammatone.py:196-237, :313-316): `order` passes of the
 * complex one-pole  y = z + g*x ; z = c*y  per band, gain in the first pass only.
 *   x [C][ldx] real float64; y [C][B][ldy] complex128 interleaved (ldy in complex elements);
 *   coef HOST [B][3] = (gain, Re c, Im c) with c = -a[1] of the reference's (b, a);
 *   states [C][B][order][2] complex, scipy's zi convention (c * last output of the stage), in/out. */
PFB_API int pfb_fos_bank_c128(const double *x, int64_t ldx, double *y, int64_t ldy, const double *coef, int B,
                              int order, double *states, int64_t n, int C, void *stream);

/* GammatoneFilterbank.synthesize / delay (gammatone.py:323-340), fused: out[c][t] = sum over bands (ascending) of the
 * band's gain-weighted real part delayed by its delay (first samples from delay_memory, which is then refilled with
 * the block's last samples); a band with delay 0 contributes Re(band) * phase factor instead.
 *   bands [C][B][ldb] complex128; out [C][ldo] complex128; delay_memory [C][B][dmax] real in/out;
 *   par HOST [B][4] = gain, delay_samples, Re phase factor, Im phase factor; every delay <= min(dmax, n). */
PFB_API int pfb_gt_synthesize_c128(const double *bands, int64_t ldb, double *out, int64_t ldo, double *delay_memory,
                                   const double *par, int B, int dmax, int64_t n, int C, void *stream);

/* splweighting.weight_signal (splweighting.py:79-80): y = lfilter(b, a, x) per row, direct form II
 * transposed in scipy's operation order; b, a HOST arrays (at most 8 taps); z [R][max(nb,na)-1] delay
 * values in/out (zeros for the reference's behaviour). */
PFB_API int pfb_tf_filter_f64(const double *x, int64_t ldx, double *y, int64_t ldy, const double *b, int nb,
                              const double *a, int na, double *z, int64_t n, int R, void *stream);

PFB_API const char *pfb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PFB_SOS_H */
