/*
 * oracle/sos_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's cascaded Direct-Form-II biquad filtering
 * (pyfilterbank/sosfilt.c) plus the two scipy-computed recurrences the reference
 * reaches through scipy.signal.lfilter (gammatone one-pole cascade, SPL weighting
 * transfer function).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object.
 *
 * Pinning: this restatement is checked bit-for-bit against the unmodified reference
 * C file compiled into oracle/_ref/libsosfilt_ref.so (tests/test_oracle.py) and
 * against golden vectors produced by importing the Python reference
 * (tests/golden/make_golden.py).
 *
 * Build: gcc -std=c99 -O2 -ffp-contract=off -fPIC -shared   (see oracle/Makefile)
 * -ffp-contract=off keeps every multiply and add separately rounded, which is what
 * the reference binary does on baseline x86-64 (no FMA contraction).
 *
 * Differences from the reference that are deliberate:
 *   - 64-bit sample counts / index arithmetic (the reference's `int` indexing
 *     overflows above 2^31 elements, sosfilt.c:62,72);
 *   - a non-replicated MIMO entry (`oracle_sos_bank_*`) that reads each channel
 *     once and writes [C][B][N]; arithmetic per (channel, band) is identical.
 */
#include <stddef.h>
#include <stdint.h>

/* One cascade, in place.  Follows sosfilt.c:31-54 (double) / :5-28 (float):
 * section-outer, sample-inner; w0 = x - a1*w1 - a2*w2 evaluated left to right;
 * y = b0*w0 + b1*w1 + b2*w2 evaluated left to right; a0 (sos[3]) is skipped. */
#define DEFINE_CASCADE(NAME, T)                                                   \
    static void NAME(T *sig, int64_t n, const T *sos, int ksos, T *st)            \
    {                                                                             \
        for (int k = 0; k < ksos; ++k) {                                          \
            const T *q = sos + (size_t)6 * k;                                     \
            const T b0 = q[0], b1 = q[1], b2 = q[2], a1 = q[4], a2 = q[5];        \
            T w1 = st[2 * k], w2 = st[2 * k + 1];                                 \
            for (int64_t i = 0; i < n; ++i) {                                     \
                T w0 = sig[i];                                                    \
                w0 = w0 - a1 * w1 - a2 * w2;                                      \
                T y = b0 * w0 + b1 * w1 + b2 * w2;                                \
                w2 = w1;                                                          \
                w1 = w0;                                                          \
                sig[i] = y;                                                       \
            }                                                                     \
            st[2 * k] = w1;                                                       \
            st[2 * k + 1] = w2;                                                   \
        }                                                                         \
    }

DEFINE_CASCADE(cascade_f32, float)
DEFINE_CASCADE(cascade_f64, double)

/* sosfilt.c:5-28 */
void oracle_sosfilter(float *signal, int64_t nsamp, const float *sos, int ksos, float *states)
{
    cascade_f32(signal, nsamp, sos, ksos, states);
}

/* sosfilt.c:31-54 */
void oracle_sosfilter_double(double *signal, int64_t nsamp, const double *sos, int ksos,
                             double *states)
{
    cascade_f64(signal, nsamp, sos, ksos, states);
}

/* sosfilt.c:57-84: signal [C][B][N] pre-filled by the caller, sos consumed
 * sequentially over (c, b, k) => [C][B][K][6], states [C][B][K][2]. */
void oracle_sosfilter_double_mimo(double *signal, int64_t nframes, int nchan, const double *sos,
                                  int ksos, int kbands, double *states)
{
    for (int c = 0; c < nchan; ++c)
        for (int b = 0; b < kbands; ++b) {
            size_t row = (size_t)c * kbands + b;
            cascade_f64(signal + row * (size_t)nframes, nframes, sos + row * 6 * (size_t)ksos,
                        ksos, states + row * 2 * (size_t)ksos);
        }
}

/* Bank form used by the parity tests: x [C][ldx] is read once per (c, b), y is
 * [C][B][ldy].  sos is [B][K][6] (sos_per_channel == 0) or [C][B][K][6].
 * Per (c, b) the arithmetic is exactly cascade_*(). */
#define DEFINE_BANK(NAME, T, CASCADE)                                                       \
    void NAME(const T *x, int64_t ldx, T *y, int64_t ldy, int64_t n, int nchan, int kbands, \
              int ksos, const T *sos, int sos_per_channel, T *states)                       \
    {                                                                                       \
        for (int c = 0; c < nchan; ++c)                                                     \
            for (int b = 0; b < kbands; ++b) {                                              \
                size_t row = (size_t)c * kbands + b;                                        \
                size_t srow = sos_per_channel ? row : (size_t)b;                            \
                T *out = y + row * (size_t)ldy;                                             \
                const T *in = x + (size_t)c * (size_t)ldx;                                  \
                for (int64_t i = 0; i < n; ++i) out[i] = in[i];                             \
                CASCADE(out, n, sos + srow * 6 * (size_t)ksos, ksos,                        \
                        states + row * 2 * (size_t)ksos);                                   \
            }                                                                               \
    }

DEFINE_BANK(oracle_sos_bank_f32, float, cascade_f32)
DEFINE_BANK(oracle_sos_bank_f64, double, cascade_f64)

/* Complex one-pole cascade of GammatoneFilterbank.analyze (gammatone.py:196-237):
 * `order` passes of scipy.signal.lfilter(b=[g], a=[1,-c]) with zi=[z]; the gain g is
 * applied in the first pass only (gammatone.py:236).  scipy's lfilter is direct form II
 * transposed; for these lengths it reduces to  y = z + g*x ; z = c*y  (a[1] = -c, so
 * -a[1]*y = c*y).  Complex products are written out as scipy's C loops do
 * (re = ar*br - ai*bi, im = ar*bi + ai*br).  State convention: z = c*y[last].
 * x real [n]; out interleaved complex [n][2]; coef = (cr, ci); states [order][2] in/out. */
void oracle_fos_cascade_c128(const double *x, int64_t n, double gain, double cr, double ci,
                             int order, double *out, double *states)
{
    for (int64_t i = 0; i < n; ++i) {
        out[2 * i] = x[i];
        out[2 * i + 1] = 0.0;
    }
    for (int s = 0; s < order; ++s) {
        double g = (s == 0) ? gain : 1.0;
        double zr = states[2 * s], zi = states[2 * s + 1];
        for (int64_t i = 0; i < n; ++i) {
            double xr = out[2 * i], xi = out[2 * i + 1];
            double yr = zr + g * xr;
            double yi = zi + g * xi;
            zr = cr * yr - ci * yi;
            zi = cr * yi + ci * yr;
            out[2 * i] = yr;
            out[2 * i + 1] = yi;
        }
        states[2 * s] = zr;
        states[2 * s + 1] = zi;
    }
}

/* SPL weighting (splweighting.py:79-80): y = lfilter(b, a, x) with a[0] == 1 after
 * scipy's normalisation, zero initial state.  Direct form II transposed in scipy's
 * operation order (_linear_filter in scipy/signal/_lfilter.c.in):
 *   y    = z[0] + b[0]*x
 *   z[k] = z[k+1] + x*b[k+1] - y*a[k+1]     k = 0 .. nt-3
 *   z[nt-2] = x*b[nt-1] - y*a[nt-1]
 * b and a have `ntaps` entries each (zero padded by the caller).  z [ntaps-1] in/out. */
void oracle_tf_df2t_f64(const double *x, double *y, int64_t n, const double *b, const double *a,
                        int ntaps, double *z)
{
    for (int64_t i = 0; i < n; ++i) {
        double xi = x[i];
        double yi;
        if (ntaps > 1) {
            yi = z[0] + b[0] * xi;
            for (int k = 0; k < ntaps - 2; ++k) z[k] = z[k + 1] + xi * b[k + 1] - yi * a[k + 1];
            z[ntaps - 2] = xi * b[ntaps - 1] - yi * a[ntaps - 1];
        } else {
            yi = b[0] * xi;
        }
        y[i] = yi;
    }
}
