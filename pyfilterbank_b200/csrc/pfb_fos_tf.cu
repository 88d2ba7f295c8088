// pfb_fos_tf.cu -- the two scipy.signal.lfilter recurrences the reference reaches from its hot path:
//
//   fos_seq_kernel  complex one-pole cascade of GammatoneFilterbank.analyze / fosfilter
//                   (pyfilterbank/gammatone.py:196-237, :313-316): `order` passes of
//                   y = z + g*x ; z = c*y  (scipy's direct form II transposed for b=[g], a=[1,-c]),
//                   gain g in the first pass only.  Real float64 input [C][N], complex128 output
//                   [C][B][N] (interleaved re, im), states [C][B][order] complex = c*y[last].
//   tf_seq_kernel   transfer-function filter of splweighting.weight_signal (splweighting.py:79-80):
//                   y = lfilter(b, a, x), direct form II transposed in scipy's operation order, up to
//                   8 taps (A-weighting has 7), one row per lane.
//
// One lane per cascade / row, sequential in time, through the same swizzled row-tile engine as the
// SOS kernels (pfb_kernels.cuh).
#include <cstdlib>

#include "pfb_kernels.cuh"
#include "../../include/pfb_sos.h"

namespace pfb {

struct FosParams {
    const double *x;
    double *y;            // complex interleaved
    const double *coef;   // device [B][3] = gain, cr, ci
    double *states;       // [Q][order][2]
    long long n, ldx, ldy;   // ldy in complex elements
    int Q, B;
};

template <int ORDER, bool ALIGNED>
__global__ void __launch_bounds__(128, 3) fos_seq_kernel(const FosParams p) {
    extern __shared__ __align__(128) char smem[];
    constexpr int TLEN = 16;   // real input samples per tile row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, key = lane & 7;
    char *xbuf = smem + warp * 4 * kTileBytes;        // 2 input tiles, then 2 output half tiles
    char *ybuf = xbuf + 2 * kTileBytes;
    const long long q0 = (long long)blockIdx.x * blockDim.x + warp * 32;
    const long long q = q0 + lane;
    const int Q = p.Q;
    if (q0 >= Q) return;
    const bool active = q < Q;
    const long long n = p.n;

    double g = 0, cr = 0, ci = 0, zr[ORDER], zi[ORDER];
    double *st = p.states + (size_t)(active ? q : 0) * ORDER * 2;
#pragma unroll
    for (int s = 0; s < ORDER; ++s) zr[s] = zi[s] = 0;
    if (active) {
        const double *cf = p.coef + (size_t)(q % p.B) * 3;
        g = cf[0], cr = cf[1], ci = cf[2];
#pragma unroll
        for (int s = 0; s < ORDER; ++s) { zr[s] = st[2 * s]; zi[s] = st[2 * s + 1]; }
    }
    const double *xrow = p.x + (active ? q / p.B : 0) * p.ldx;
    const long long mylen = active ? n : 0;
    // output rows seen as doubles: row stride 2*ldy, row length 2*n
    CoopIO<double, ALIGNED> io(p.x, p.y + q0 * 2 * p.ldy, 0, 2 * p.ldy, lane);
    auto rowlen2 = [=](int row) -> long long { return (q0 + row < Q) ? 2 * n : 0; };
    char *myx = xbuf + lane * kRowBytes;
    const bool warp_full = q0 + 32 <= Q;

    const long long nt = (n + TLEN - 1) / TLEN;
    if (nt > 0) {
        load_row_own<double, ALIGNED>(myx, key, xrow, mylen, 0, p.x);
        cp_async_commit();
    }
    for (long long t = 0; t < nt; ++t) {
        const int cb = (int)(t & 1);
        if (t + 1 < nt) {
            load_row_own<double, ALIGNED>(myx + (cb ^ 1) * kTileBytes, key, xrow, mylen, (t + 1) * TLEN, p.x);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        const int valid = clamp_i(mylen - t * TLEN, TLEN);
        const char *xr_row = myx + cb * kTileBytes;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const double2 in = *reinterpret_cast<const double2 *>(xr_row + ((v ^ key) << 4));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int smp = 2 * v + i;
                if (smp < valid) {
                    double yr = i ? in.y : in.x, yi = 0.0;
#pragma unroll
                    for (int s = 0; s < ORDER; ++s) {
                        const double gg = s == 0 ? g : 1.0;
                        yr = zr[s] + gg * yr;
                        yi = zi[s] + gg * yi;
                        zr[s] = cr * yr - ci * yi;
                        zi[s] = cr * yi + ci * yr;
                    }
                    // complex sample smp of the tile -> half tile smp / 8, 16-byte chunk smp % 8
                    *reinterpret_cast<double2 *>(ybuf + (smp >> 3) * kTileBytes + lane * kRowBytes +
                                                 (((smp & 7) ^ key) << 4)) = make_double2(yr, yi);
                }
            }
        }
        __syncwarp();
        const bool interior = (t + 1) * TLEN <= n;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long e0 = 2 * (t * TLEN + 8 * h);
            if (interior && warp_full) io.store_fast(ybuf + h * kTileBytes, e0);
            else io.store(ybuf + h * kTileBytes, e0, rowlen2);
        }
        __syncwarp();
    }
    if (active && n > 0) {
#pragma unroll
        for (int s = 0; s < ORDER; ++s) { st[2 * s] = zr[s]; st[2 * s + 1] = zi[s]; }
    }
}

struct TfParams {
    const double *x;
    double *y;
    double *z;            // [R][ntaps-1] in/out
    long long n, ldx, ldy;
    int R, ntaps;
    double b[8], a[8];    // a[0] == 1 (normalised on the host like scipy does)
};

// One scipy _linear_filter step (direct form II transposed, a[0] == 1), every product and sum rounded separately
// in scipy's order: the weighting filters have a 4-fold zero at z = 1, so FMA contraction alone moves the
// result by ~1e-9 relative.   y = z0 + b0 x ; z[k] = (z[k+1] + x b[k+1]) - y a[k+1] ; z[NZ-1] = x b[NZ] - y a[NZ]
template <int NZ>
__device__ __forceinline__ double tf_step(double xi, const double (&b)[8], const double (&a)[8], double (&z)[NZ]) {
    const double yi = __dadd_rn(z[0], __dmul_rn(b[0], xi));
#pragma unroll
    for (int k = 0; k < NZ - 1; ++k)
        z[k] = __dsub_rn(__dadd_rn(z[k + 1], __dmul_rn(xi, b[k + 1])), __dmul_rn(yi, a[k + 1]));
    z[NZ - 1] = __dsub_rn(__dmul_rn(xi, b[NZ]), __dmul_rn(yi, a[NZ]));
    return yi;
}

// One lane per row, one warp per CTA (the recurrence is latency bound: 3 dependent FP64 operations per sample,
// so rows are spread over as many SM sub-partitions as possible).  NZ = taps - 1 delay values.
template <int NZ, bool ALIGNED, int kTfStages>
__global__ void __launch_bounds__(32) tf_seq_kernel(const __grid_constant__ TfParams p) {
    extern __shared__ __align__(128) char smem[];
    constexpr int TLEN = 16;
    const int lane = threadIdx.x & 31, key = lane & 7;
    char *buf0 = smem;
    const long long r0 = (long long)blockIdx.x * 32;
    const long long r = r0 + lane;
    const int R = p.R;
    if (r0 >= R) return;
    const bool active = r < R;
    const long long n = p.n;
    double z[NZ];
#pragma unroll
    for (int k = 0; k < NZ; ++k) z[k] = active ? p.z[(size_t)r * NZ + k] : 0.0;
    double b[8], a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { b[k] = p.b[k]; a[k] = p.a[k]; }
    const double *xrow = p.x + (active ? r : 0) * p.ldx;
    const long long mylen = active ? n : 0;
    CoopIO<double, ALIGNED> io(p.x, p.y + r0 * p.ldy, 0, p.ldy, lane);
    auto rowlen = [=](int row) -> long long { return (r0 + row < R) ? n : 0; };
    char *myrow = buf0 + lane * kRowBytes;
    const bool warp_full = r0 + 32 <= R;
    const long long nt = (n + TLEN - 1) / TLEN;
    // kTfStages tiles in flight: one tile is only ~16 x 50 clocks of arithmetic, less than a DRAM round trip
#pragma unroll
    for (int t = 0; t < kTfStages - 1; ++t) {
        if (t < nt) load_row_own<double, ALIGNED>(myrow + t * kTileBytes, key, xrow, mylen, (long long)t * TLEN, p.x);
        cp_async_commit();
    }
    for (long long t = 0; t < nt; ++t) {
        const int cb = (int)(t % kTfStages);
        if (t + kTfStages - 1 < nt)
            load_row_own<double, ALIGNED>(myrow + (int)((t + kTfStages - 1) % kTfStages) * kTileBytes, key, xrow, mylen,
                                          (t + kTfStages - 1) * TLEN, p.x);
        cp_async_commit();
        cp_async_wait<kTfStages - 1>();
        const int valid = clamp_i(mylen - t * TLEN, TLEN);
        char *row = myrow + cb * kTileBytes;
        if (valid == TLEN) {   // whole tile: registers in, registers out, no per-sample predicates
            double2 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const double2 *>(row + ((i ^ key) << 4));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i].x = tf_step<NZ>(v[i].x, b, a, z);
                v[i].y = tf_step<NZ>(v[i].y, b, a, z);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<double2 *>(row + ((i ^ key) << 4)) = v[i];
        } else {
            for (int i = 0; i < valid; ++i) {
                double *e = reinterpret_cast<double *>(row + (((i >> 1) ^ key) << 4) + (i & 1) * 8);
                *e = tf_step<NZ>(*e, b, a, z);
            }
        }
        __syncwarp();
        if ((t + 1) * TLEN <= n && warp_full) io.store_fast(buf0 + cb * kTileBytes, t * TLEN);
        else io.store(buf0 + cb * kTileBytes, t * TLEN, rowlen);
        __syncwarp();
    }
    if (active && n > 0) {
#pragma unroll
        for (int k = 0; k < NZ; ++k) p.z[(size_t)r * NZ + k] = z[k];
    }
}

template <int ORDER>
static cudaError_t run_fos(const FosParams &p, bool aligned, cudaStream_t s) {
    const int warps = 4, smem = warps * 4 * kTileBytes;
    const unsigned grid = (unsigned)((p.Q + warps * 32 - 1) / (warps * 32));
    cudaError_t e;
    if (aligned) {
        e = cudaFuncSetAttribute(fos_seq_kernel<ORDER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        fos_seq_kernel<ORDER, true><<<grid, warps * 32, smem, s>>>(p);
    } else {
        e = cudaFuncSetAttribute(fos_seq_kernel<ORDER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        fos_seq_kernel<ORDER, false><<<grid, warps * 32, smem, s>>>(p);
    }
    return cudaGetLastError();
}

int launch_fos(const FosParams &p, int order, bool aligned, cudaStream_t s) {
    switch (order) {
        case 1: return (int)run_fos<1>(p, aligned, s);
        case 2: return (int)run_fos<2>(p, aligned, s);
        case 3: return (int)run_fos<3>(p, aligned, s);
        case 4: return (int)run_fos<4>(p, aligned, s);
        case 5: return (int)run_fos<5>(p, aligned, s);
        case 6: return (int)run_fos<6>(p, aligned, s);
        case 8: return (int)run_fos<8>(p, aligned, s);
        default: return (int)cudaErrorInvalidValue;
    }
}

template <int NZ>
static int run_tf(const TfParams &p, bool aligned, cudaStream_t s) {
    const unsigned grid = (unsigned)((p.R + 31) / 32);
    const char *v = getenv("PFB_TF_STAGES");
    const int st = v ? atoi(v) : 2;
    if (!aligned) tf_seq_kernel<NZ, false, 2><<<grid, 32, 2 * kTileBytes, s>>>(p);
    else if (st == 4) tf_seq_kernel<NZ, true, 4><<<grid, 32, 4 * kTileBytes, s>>>(p);
    else if (st == 3) tf_seq_kernel<NZ, true, 3><<<grid, 32, 3 * kTileBytes, s>>>(p);
    else tf_seq_kernel<NZ, true, 2><<<grid, 32, 2 * kTileBytes, s>>>(p);
    return (int)cudaGetLastError();
}

int launch_tf(const TfParams &p, bool aligned, cudaStream_t s) {
    switch (p.ntaps - 1) {
        case 0:   // pure gain: one (always zero) delay value keeps the state layout [R][max(taps - 1, 1)]
        case 1: return run_tf<1>(p, aligned, s);
        case 2: return run_tf<2>(p, aligned, s);
        case 3: return run_tf<3>(p, aligned, s);
        case 4: return run_tf<4>(p, aligned, s);
        case 5: return run_tf<5>(p, aligned, s);
        case 6: return run_tf<6>(p, aligned, s);
        case 7: return run_tf<7>(p, aligned, s);
        default: return (int)cudaErrorInvalidValue;
    }
}

}  // namespace pfb

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern int pfb_set_error(int status, const char *msg);   // pfb_api.cu

extern "C" {

int pfb_fos_bank_c128(const double *x, int64_t ldx, double *y, int64_t ldy, const double *coef_host, int B,
                      int order, double *states, int64_t n, int C, void *stream) {
    if (n == 0) return pfb_set_error(PFB_OK, "");
    if (!x || !y || !coef_host || !states || B <= 0 || C <= 0 || n < 0 || ldx < n || ldy < n)
        return pfb_set_error(PFB_ERR_INVALID, "pfb_fos_bank_c128: bad argument");
    if (order < 1 || order == 7 || order > 8)
        return pfb_set_error(PFB_ERR_INVALID, "pfb_fos_bank_c128: order must be 1..6 or 8");
    if ((long long)C * B > 0x7fffffffLL) return pfb_set_error(PFB_ERR_INVALID, "pfb_fos_bank_c128: C*B too large");
    cudaStream_t s = (cudaStream_t)stream;
    double *dcoef = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&dcoef, (size_t)B * 3 * sizeof(double), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dcoef, coef_host, (size_t)B * 3 * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return pfb_set_error(PFB_ERR_CUDA, cudaGetErrorString(e));
    pfb::FosParams p{x, y, dcoef, states, n, ldx, ldy, C * B, B};
    const bool aligned = (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && (ldx * 8) % 16 == 0;
    int err = pfb::launch_fos(p, order, aligned, s);
    cudaFreeAsync(dcoef, s);
    if (err) return pfb_set_error(PFB_ERR_CUDA, cudaGetErrorString((cudaError_t)err));
    return pfb_set_error(PFB_OK, "");
}

int pfb_tf_filter_f64(const double *x, int64_t ldx, double *y, int64_t ldy, const double *b, int nb,
                      const double *a, int na, double *z, int64_t n, int R, void *stream) {
    if (n == 0) return pfb_set_error(PFB_OK, "");
    if (!x || !y || !b || !a || !z || nb < 1 || na < 1 || nb > 8 || na > 8 || R <= 0 || n < 0 || ldx < n || ldy < n)
        return pfb_set_error(PFB_ERR_INVALID, "pfb_tf_filter_f64: bad argument (at most 8 taps)");
    if (a[0] == 0.0) return pfb_set_error(PFB_ERR_INVALID, "pfb_tf_filter_f64: a[0] must not be 0");
    pfb::TfParams p{};
    p.x = x, p.y = y, p.z = z, p.n = n, p.ldx = ldx, p.ldy = ldy, p.R = R;
    p.ntaps = nb > na ? nb : na;
    for (int k = 0; k < 8; ++k) {   // scipy normalises by a[0] first
        p.b[k] = k < nb ? b[k] / a[0] : 0.0;
        p.a[k] = k < na ? a[k] / a[0] : 0.0;
    }
    const bool aligned = (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && (ldx * 8) % 16 == 0 && (ldy * 8) % 16 == 0;
    int err = pfb::launch_tf(p, aligned, (cudaStream_t)stream);
    if (err) return pfb_set_error(PFB_ERR_CUDA, cudaGetErrorString((cudaError_t)err));
    return pfb_set_error(PFB_OK, "");
}

}  // extern "C"
