// pfb_design.cu -- on-device coefficient design and the Direct-Form-II-Transposed state adapter
// (SURVEY.md section 8, row f4).
//
//   * pfb_design_butter_sos: digital Butterworth cascades (band-pass, band-stop, low-pass, high-pass) for B bands
//     at once, one thread per band -- the arithmetic of the reference's butter_sos / butter_analog_sos
//     (pyfilterbank/butterworth.py:16-146) and bilinear_sos (pyfilterbank/sosfiltering.py:382-432): analog
//     one-pole sections (d0 s + d1) / (c0 s + c1) with complex weights, pole splitting for the band transforms,
//     squared magnitudes under the bilinear transform.  Rows come out as the reference returns them
//     ([b0 b1 b2 a0 a1 a2], a0 == 1, sections reversed: butterworth.py:57), i.e. column b of
//     FractionalOctaveFilterbank.sosmat (octbank.py:165-172).  Matters when banks are re-designed per call (the
//     property setters of octbank.py:286-347 re-design on every change).
//   * pfb_states_convert: states of the reference's 'py' filter function (sosfilter_py, sosfiltering.py:368-379:
//     scipy.signal.lfilter per section, zi = DF-II-T delay values) <-> the (w1, w2) DF-II delay values that
//     sosfilt.c and the kernels here carry.  Both describe the same filter memory:
//         z1 = (b1 - a1 b0) w1 + (b2 - a2 b0) w2,    z2 = (b2 - a2 b0) w1 + (b2 a1 - a2 b1) w2.
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/pfb_sos.h"

int pfb_set_error(int status, const char *msg);   // pfb_api.cu

namespace {

struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx C(double re, double im = 0.0) { return cplx{re, im}; }
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return C(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return C(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx operator-(cplx a) { return C(-a.re, -a.im); }
__device__ __forceinline__ cplx operator*(cplx a, cplx b) { return C(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ cplx operator*(double s, cplx a) { return C(s * a.re, s * a.im); }
__device__ __forceinline__ cplx operator/(cplx a, cplx b) {
    // Smith's method, like numpy's complex division
    if (fabs(b.re) >= fabs(b.im)) {
        const double r = b.im / b.re, den = b.re + b.im * r;
        return C((a.re + a.im * r) / den, (a.im - a.re * r) / den);
    }
    const double r = b.re / b.im, den = b.re * r + b.im;
    return C((a.re * r + a.im) / den, (a.im * r - a.re) / den);
}
__device__ __forceinline__ cplx conj(cplx a) { return C(a.re, -a.im); }
__device__ __forceinline__ double abs2(cplx a) {   // |a|^2 the way numpy computes abs(a)**2: hypot, then square
    const double h = hypot(a.re, a.im);
    return h * h;
}
__device__ cplx csqrt(cplx z) {   // principal branch (C99 csqrt / numpy.sqrt)
    if (z.re == 0.0 && z.im == 0.0) return C(0.0, z.im);
    const double m = hypot(z.re, z.im);
    if (z.re >= 0.0) {
        const double t = sqrt(0.5 * (m + z.re));
        return C(t, z.im / (2.0 * t));
    }
    const double t = sqrt(0.5 * (m - z.re));
    return C(fabs(z.im) / (2.0 * t), copysign(t, z.im));
}

// One digital section from one analog one-pole section: bilinear_sos (sosfiltering.py:412-431), row = [b a] / a0
__device__ void bilinear_row(cplx d0, cplx d1, cplx c0, cplx c1, double *row) {
    const double a0 = abs2(c0 + c1);
    const double a1 = 2.0 * ((c0 + c1) * conj(c1 - c0)).re;
    const double a2 = abs2(c1 - c0);
    const double b0 = abs2(d0 + d1);
    const double b1 = 2.0 * ((d0 + d1) * conj(d1 - d0)).re;
    const double b2 = abs2(d1 - d0);
    row[0] = b0 / a0, row[1] = b1 / a0, row[2] = b2 / a0;
    row[3] = a0 / a0, row[4] = a1 / a0, row[5] = a2 / a0;
}

// kind: 0 low-pass, 1 high-pass, 2 band-pass, 3 band-stop.  L: low-pass poles (even).  v1[b], v2[b]: critical
// frequencies in cycles per sample.  sos [B][K][6], K = L / 2 (low / high-pass) or L (band-pass / -stop).
__global__ void butter_sos_kernel(int kind, int L, const double *v1, const double *v2, int B, double *sos, int *bad) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int half = L / 2;
    const int K = kind >= 2 ? L : half;
    double *out = sos + (size_t)b * K * 6;
    const double f1 = v1[b], f2 = kind >= 2 ? v2[b] : 0.0;
    if (!(f1 > 0.0 && f1 < 0.5) || (kind >= 2 && !(f2 > f1 && f2 < 0.5))) {   // butterworth.py:44-47
        atomicExch(bad, b + 1);
        for (int i = 0; i < K * 6; ++i) out[i] = 0.0;
        return;
    }
    const double kPi = 3.141592653589793;
    const double w1 = tan(kPi * f1), w2 = kind >= 2 ? tan(kPi * f2) : 0.0;
    const double wc = kind >= 2 ? w2 - w1 : w1;
    for (int k = 1; k <= half; ++k) {
        // low-pass prototype section wc / (s - p), p = wc exp(-j (2k + L - 1) pi / (2L))   (butterworth.py:96-102)
        const double ang = -(2.0 * k + L - 1.0) * kPi / (2.0 * L);
        double sn, cs;
        sincos(ang, &sn, &cs);
        const cplx p = C(wc * cs, wc * sn);
        cplx d0 = C(0.0), d1 = C(wc), c0 = C(1.0), c1 = -p;
        // the reference returns the sections reversed (flipud, butterworth.py:57): analog section i -> row K - 1 - i
        if (kind == 0) {
            bilinear_row(d0, d1, c0, c1, out + (size_t)(K - k) * 6);
        } else if (kind == 1) {   // s -> wc^2 / s   (butterworth.py:104-108)
            bilinear_row(d1, C(0.0), c1, wc * wc * c0, out + (size_t)(K - k) * 6);
        } else if (kind == 2) {   // s -> (s^2 + w1 w2) / s: every pole splits in two   (butterworth.py:110-124)
            const cplx root = csqrt(c1 * c1 - (4.0 * w1 * w2) * (c0 * c0));
            const cplx r1 = (-c1 + root) / (2.0 * c0), r2 = (-c1 - root) / (2.0 * c0);
            bilinear_row(d1, C(0.0), c0, -(c0 * r1), out + (size_t)(K - k) * 6);                 // analog row k - 1
            bilinear_row(C(0.0), C(1.0), C(1.0), -r2, out + (size_t)(K - half - k) * 6);         // analog row half + k - 1
        } else {                  // s -> wc^2 s / (s^2 + w1 w2)   (butterworth.py:126-144)
            const double wc2 = wc * wc, wc4 = wc2 * wc2;
            const cplx rd = csqrt(wc4 * (d0 * d0) - (4.0 * w1 * w2) * (d1 * d1));
            const cplx dr1 = (-(wc2 * d0) + rd) / (2.0 * d1), dr2 = (-(wc2 * d0) - rd) / (2.0 * d1);
            const cplx rc = csqrt(wc4 * (c0 * c0) - (4.0 * w1 * w2) * (c1 * c1));
            const cplx cr1 = (-(wc2 * c0) + rc) / (2.0 * c1), cr2 = (-(wc2 * c0) - rc) / (2.0 * c1);
            bilinear_row(d1, -(d1 * dr1), c1, -(c1 * cr1), out + (size_t)(K - k) * 6);
            bilinear_row(C(1.0), -dr2, C(1.0), -cr2, out + (size_t)(K - half - k) * 6);
        }
    }
}

// states [Q][K][2] in place; sos [B][K][6] (shared) or [Q][K][6]; to_df2t: (w1, w2) -> (z1, z2), else the inverse
__global__ void states_convert_kernel(double *states, const double *sos, long long Q, int B, int K, int per_row, int to_df2t,
                                      int *bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (q, k)
    if (i >= Q * K) return;
    const long long q = i / K;
    const int k = (int)(i - q * K);
    const double *c = sos + ((size_t)(per_row ? q : q % B) * K + k) * 6;
    const double a0 = c[3] != 0.0 ? c[3] : 1.0;                              // scipy normalises by a0; sosfilt.c ignores it
    const double b0 = c[0] / a0, b1 = c[1] / a0, b2 = c[2] / a0, a1 = c[4] / a0, a2 = c[5] / a0;
    const double t11 = b1 - a1 * b0, t12 = b2 - a2 * b0, t22 = b2 * a1 - a2 * b1;   // T = [[t11 t12] [t12 t22]]
    double *s = states + (size_t)i * 2;
    const double u = s[0], v = s[1];
    if (to_df2t) {
        s[0] = t11 * u + t12 * v;
        s[1] = t12 * u + t22 * v;
    } else {
        const double det = t11 * t22 - t12 * t12;
        if (det == 0.0) {          // numerator and denominator share a root: the section's memory is not observable
            atomicExch(bad, 1);
            return;
        }
        s[0] = (t22 * u - t12 * v) / det;
        s[1] = (t11 * v - t12 * u) / det;
    }
}

int fail(int status, const char *msg) { return pfb_set_error(status, msg); }

}  // namespace

extern "C" {

int pfb_design_butter_sos(int kind, int L, const double *v1, const double *v2, int B, double *sos_host, double *sos_dev,
                          void *stream_) {
    if (kind < 0 || kind > 3 || L < 2 || (L & 1) || L > 64 || !v1 || (kind >= 2 && !v2) || B <= 0 || (!sos_host && !sos_dev))
        return fail(PFB_ERR_INVALID, "pfb_design_butter_sos: bad argument (L must be even, kind 0..3)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PFB_ERR_NO_DEVICE, "pfb_design_butter_sos: no CUDA device");
    cudaStream_t s = (cudaStream_t)stream_;
    const int K = kind >= 2 ? L : L / 2;
    const size_t nv = (size_t)B * sizeof(double), nsos = (size_t)B * K * 6 * sizeof(double);
    char *scratch = nullptr;     // [v1][v2][bad][sos when the caller gave no device buffer]
    const size_t off_bad = 2 * nv, off_sos = (off_bad + 256) / 256 * 256;
    cudaError_t e = cudaMallocAsync((void **)&scratch, off_sos + (sos_dev ? 0 : nsos), s);
    if (e != cudaSuccess) return fail(PFB_ERR_ALLOC, cudaGetErrorString(e));
    double *dv1 = (double *)scratch, *dv2 = (double *)(scratch + nv);
    int *dbad = (int *)(scratch + off_bad);
    double *dsos = sos_dev ? sos_dev : (double *)(scratch + off_sos);
    e = cudaMemcpyAsync(dv1, v1, nv, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && kind >= 2) e = cudaMemcpyAsync(dv2, v2, nv, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(dbad, 0, sizeof(int), s);
    if (e == cudaSuccess) {
        butter_sos_kernel<<<(B + 63) / 64, 64, 0, s>>>(kind, L, dv1, dv2, B, dsos, dbad);
        e = cudaGetLastError();
    }
    int bad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && sos_host) e = cudaMemcpyAsync(sos_host, dsos, nsos, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFreeAsync(scratch, s);
    if (e != cudaSuccess) return fail(PFB_ERR_CUDA, cudaGetErrorString(e));
    if (bad) {
        char msg[160];
        snprintf(msg, sizeof msg, "pfb_design_butter_sos: band %d needs 0 < v1 < v2 < 0.5 (butterworth.py:44-47)", bad - 1);
        return fail(PFB_ERR_INVALID, msg);
    }
    return pfb_set_error(PFB_OK, "");
}

int pfb_states_convert(double *states_dev, const double *sos, int64_t Q, int B, int K, int per_row, int to_df2t, void *stream_) {
    if (!states_dev || !sos || Q <= 0 || B <= 0 || K <= 0) return fail(PFB_ERR_INVALID, "pfb_states_convert: bad argument");
    cudaStream_t s = (cudaStream_t)stream_;
    const size_t nsos = (size_t)(per_row ? Q : B) * K * 6 * sizeof(double);
    char *scratch = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&scratch, nsos + 256, s);
    if (e != cudaSuccess) return fail(PFB_ERR_ALLOC, cudaGetErrorString(e));
    int *dbad = (int *)(scratch + nsos);
    e = cudaMemcpyAsync(scratch, sos, nsos, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(dbad, 0, sizeof(int), s);
    if (e == cudaSuccess) {
        const long long n = Q * K;
        states_convert_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(states_dev, (const double *)scratch, Q, B, K, per_row,
                                                                         to_df2t, dbad);
        e = cudaGetLastError();
    }
    int bad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFreeAsync(scratch, s);
    if (e != cudaSuccess) return fail(PFB_ERR_CUDA, cudaGetErrorString(e));
    if (bad) return fail(PFB_ERR_INVALID, "pfb_states_convert: a section's numerator and denominator share a root");
    return pfb_set_error(PFB_OK, "");
}

}  // extern "C"
