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
#define PFB_ARITH_AUTO 0x200 /* with PFB_ARITH_F64 (float32 signals, float64 states): the arithmetic is chosen PER BAND --
                               float32 where a calibration run at bank level shows float32 arithmetic within 1e-5
                               (relative L2) of float64, float64 elsewhere -- so every band stays within 1e-4 of the
                               double result while the well-conditioned bands run at the float32 rate
                               (pfb_bank_f32_bands() tells how many; time-chunked TMA kernels only, other paths
                               compute everything in float64) */
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
/* The same with the SPL-weighting prefilter in front (splweighting.py:61-80 `weight_signal`, then `filter_mimo_c`, then
 * 10 log10 sum y^2 -- octbank.py:262-263): prefilter, bank and energies in ONE kernel, only the input is read from HBM.
 * wb / wa: transfer-function taps like pfb_bank_filter_weighted (5 to 7 taps), wz DEVICE [C][ntaps-1] delay values
 * (scipy's zi) in/out. */
PFB_API int pfb_bank_levels_weighted(pfb_bank *bank, int dtype, const void *x, int64_t ldx, double *levels, int64_t ld_levels,
                                     void *states, const double *wb, int nwb, const double *wa, int nwa, double *wz,
                                     int64_t n, int C, int64_t block_len, int flags, void *stream);

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
                             with the 512-byte-fragment output pass; 4: TMA kernels with the fused weighting warp;
                             5: TMA kernels with a decimating band */
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

/* ------------------------------------------------------------------------------------------------
 * 4. Fused callers of the bank (a8, a6', f2).
 * ---------------------------------------------------------------------------------------------- */

/* FractionalOctaveFilterbank.filter_mimo_c on the SPL-weighted signal as ONE call: y = bank(lfilter(b, a, x)) with
 * lfilter the reference's weighting prefilter (splweighting.weight_signal, splweighting.py:61-80: direct form II
 * transposed, scipy's operation order, float64 arithmetic also for float32 signals).  With enough channels
 * (32-channel blocks x band groups fill the machine) the prefilter runs as one more warp inside the bank kernel on
 * the shared-memory input tiles and the weighted signal never touches HBM; otherwise the weighting kernel and the
 * time-chunked bank run back to back on a scratch buffer.  DEVICE pointers, asynchronous on `stream`.
 *   wb / wa : HOST taps (at most 8 each, normalised by wa[0] like scipy does)
 *   wz      : DEVICE [C][max(nwb, nwa) - 1] delay values in scipy's zi convention, in/out (zeros = the reference's call)
 * Everything else as in pfb_bank_filter (flags: mode / PFB_EXACT / PFB_ARITH_F64). */
PFB_API int pfb_bank_filter_weighted(pfb_bank *bank, int dtype, const void *x, int64_t ldx, void *y, int64_t ldy,
                                     void *states, const double *wb, int nwb, const double *wa, int nwa, double *wz,
                                     int64_t n, int C, int flags, void *stream);

/* Decimating band path (opt-in; the reference has no resampling): stage m = 0 .. nstages - 1
 * runs at fs / 2^m; the stage m + 1 signal is the stage m signal through the anti-alias cascade `aa_sos`, every
 * second sample kept.  One call filters all stages.  Default schedule ("split"): the anti-alias chain (one cascade per
 * channel and stage, stored decimated) runs on a stream of its own, stage 0's bands on the caller's stream, the bands
 * of stage m >= 1 on stream m as soon as signal m exists, and the caller's stream continues when all of them have
 * finished -- the small stages overlap each other and stage 0.  PFB_DEC_SPLIT=0 (environment) or PFB_EXACT: stage by
 * stage, per stage ONE pass over the stage signal produces the stage's band signals and (as one more band warp with a
 * decimating store) the next stage's signal.  Either way a band signal is written once, at its own rate.
 *   bands_per_stage[m], sos_per_stage[m] : HOST [bands][K][6] (null for a stage without bands)
 *   aa_sos : HOST [Kaa][6]                                                                                   */
typedef struct pfb_decbank pfb_decbank;
PFB_API int pfb_decbank_create(pfb_decbank **out, int nstages, const int *bands_per_stage,
                               const double *const *sos_per_stage, int K, const double *aa_sos, int Kaa);
PFB_API void pfb_decbank_destroy(pfb_decbank *dec);
/* x DEVICE [C][ldx] float64, n samples; y[m] DEVICE [C][bands m][ldy[m]] receives n_m samples per row with n_0 = n,
 * n_{m+1} = (n_m - phase_m + 1) / 2; band_states[m] DEVICE [C][bands m][K][2]; aa_states DEVICE
 * [nstages - 1][C][Kp][2] (Kp from pfb_decbank_info: the anti-alias cascade padded with pass-through sections);
 * phases HOST [nstages - 1] decimator phases, in/out.  flags: mode / PFB_EXACT.  Asynchronous: later work on `stream`
 * is ordered behind all of it; the buffers must not be touched from other streams before. */
PFB_API int pfb_decbank_filter(pfb_decbank *dec, const double *x, int64_t ldx, void *const *y, const int64_t *ldy,
                               void *const *band_states, double *aa_states, int *phases, int64_t n, int C, int flags,
                               void *stream);
PFB_API int pfb_decbank_info(const pfb_decbank *dec, int *stages, int *aa_sections_padded, int *last_kernels,
                             int *last_fused_stages);

/* Streaming session: block-wise filtering of HOST audio with carried, device-resident states (what the reference's
 * `states` arguments exist for: blockprocessing.py:3-72, sosfiltering.py:143-147).  Streams, events, staging slots
 * and the state array are created once; pfb_stream_push enqueues input copy, kernels and output copy of one block
 * and returns (asynchronous for page-locked host buffers; the buffers must stay valid until pfb_stream_sync).  Blocks
 * of one length allocate nothing after the first.  x HOST [C][ldx], y HOST [C][B][ldy], n <= max_block samples. */
typedef struct pfb_stream pfb_stream;
typedef struct pfb_stream_info {
    int64_t blocks, samples;
    int64_t device_allocs_total;              /* device allocations of this library since pfb_stream_create */
    int64_t device_allocs_after_first_block;  /* ... since the first pfb_stream_push returned */
} pfb_stream_info;
PFB_API int pfb_stream_create(pfb_stream **out, pfb_bank *bank, int dtype, int C, int64_t max_block, int flags);
PFB_API int pfb_stream_push(pfb_stream *s, const void *x, int64_t ldx, void *y, int64_t ldy, int64_t n);
PFB_API int pfb_stream_sync(pfb_stream *s);
/* wait until the output of push number `block` (0-based) is in host memory, later pushes may still be in flight */
PFB_API int pfb_stream_wait(pfb_stream *s, int64_t block);
/* copy the carried states [C][B][K][2] to (set = 0) or from (set = 1) a HOST array; synchronises the session */
PFB_API int pfb_stream_states(pfb_stream *s, void *states_host, int set);
PFB_API int pfb_stream_stats(pfb_stream *s, pfb_stream_info *info);
PFB_API void pfb_stream_destroy(pfb_stream *s);

/* A bank keeps chunk-state scratch per CUDA stream it has been used on; call this before destroying a stream. */
PFB_API int pfb_bank_release_stream(pfb_bank *bank, void *stream);
/* device allocations made by this library so far (diagnostics / tests) */
PFB_API long long pfb_device_allocs(void);

/* ------------------------------------------------------------------------------------------------
 * On-device coefficient design and the 'py' state adapter
 * ------------------------------------------------------------------------------------------------
 * pfb_design_butter_sos replaces, for B bands at once, butter_sos(band, L, v1, v2)
 * (pyfilterbank/butterworth.py:16-58 with butter_analog_sos :61-146 and bilinear_sos,
 * pyfilterbank/sosfiltering.py:382-432): kind 0 low-pass, 1 high-pass, 2 band-pass, 3 band-stop; L low-pass poles
 * (even); v1[B] (and v2[B] for kinds 2, 3) HOST arrays of critical frequencies in cycles per sample, 0 < v1 < v2 < 0.5
 * (PFB_ERR_INVALID otherwise, like the reference's exceptions).  Result: [B][K][6] rows [b0 b1 b2 a0 a1 a2], a0 == 1,
 * sections in the reference's (reversed) order, K = L for kinds 2, 3 and L / 2 otherwise -- column b of
 * FractionalOctaveFilterbank.sosmat is row b flattened (octbank.py:165-172).  Written to sos_host (HOST, may be null)
 * and / or sos_dev (DEVICE, may be null).  Synchronises the stream. */
PFB_API int pfb_design_butter_sos(int kind, int L, const double *v1, const double *v2, int B, double *sos_host,
                                  double *sos_dev, void *stream);
/* States of the reference's 'py' filter function (sosfilter_py, sosfiltering.py:368-379: scipy.signal.lfilter per
 * section, zi = Direct-Form-II-Transposed delay values) <-> the (w1, w2) Direct-Form-II delay values of sosfilt.c and of
 * every kernel here.  states_dev DEVICE [Q][K][2], converted in place; sos HOST [B][K][6] (row q uses bank row q % B)
 * or, per_row != 0, [Q][K][6]; to_df2t != 0: (w1, w2) -> (z1, z2), else the inverse (PFB_ERR_INVALID when a section's
 * numerator and denominator share a root).  a0 normalises the section as scipy does.  Synchronises the stream. */
PFB_API int pfb_states_convert(double *states_dev, const double *sos, int64_t Q, int B, int K, int per_row, int to_df2t,
                               void *stream);

/* number of bands PFB_ARITH_AUTO runs in float32 arithmetic (calibrated on first use; negative: error) */
PFB_API int pfb_bank_f32_bands(pfb_bank *bank);

PFB_API const char *pfb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PFB_SOS_H */
