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

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pfb_kernels.cuh"
#include "pfb_fos_tma.cuh"
#include "pfb_tf.cuh"
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

// One lane per row, one warp per CTA.  Measured (scripts/tf_time.py): a warp needs ~94 clocks per sample for the
// 25 separately rounded FP64 operations of the 7-tap filter (57 of them arithmetic, the rest tile loads and stores:
// scripts/sass_sched.py), and two rows per lane take exactly twice as long, so the limit is the single-warp issue
// rate, not the dependent chain; rows are therefore spread over as many SM sub-partitions as possible (one warp
// each).  NZ = taps - 1 delay values.  T: float64 rows, or float32 rows filtered in float64 arithmetic.
template <typename T, int NZ, bool ALIGNED, int kTfStages>
__global__ void __launch_bounds__(32) tf_seq_kernel(const __grid_constant__ TfParams p) {
    extern __shared__ __align__(128) char smem[];
    using V = typename Vec16<T>::type;
    constexpr int EPV = Vec16<T>::N;
    constexpr int TLEN = kRowBytes / (int)sizeof(T);
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
    const T *xbase = reinterpret_cast<const T *>(p.x);
    const T *xrow = xbase + (active ? r : 0) * p.ldx;
    const long long mylen = active ? n : 0;
    CoopIO<T, ALIGNED> io(xbase, reinterpret_cast<T *>(p.y) + r0 * p.ldy, 0, p.ldy, lane);
    auto rowlen = [=](int row) -> long long { return (r0 + row < R) ? n : 0; };
    char *myrow = buf0 + lane * kRowBytes;
    const bool warp_full = r0 + 32 <= R;
    const long long nt = (n + TLEN - 1) / TLEN;
    // kTfStages tiles in flight: one tile is only ~16 x 50 clocks of arithmetic, less than a DRAM round trip
#pragma unroll
    for (int t = 0; t < kTfStages - 1; ++t) {
        if (t < nt) load_row_own<T, ALIGNED>(myrow + t * kTileBytes, key, xrow, mylen, (long long)t * TLEN, xbase);
        cp_async_commit();
    }
    for (long long t = 0; t < nt; ++t) {
        const int cb = (int)(t % kTfStages);
        if (t + kTfStages - 1 < nt)
            load_row_own<T, ALIGNED>(myrow + (int)((t + kTfStages - 1) % kTfStages) * kTileBytes, key, xrow, mylen,
                                     (t + kTfStages - 1) * TLEN, xbase);
        cp_async_commit();
        cp_async_wait<kTfStages - 1>();
        const int valid = clamp_i(mylen - t * TLEN, TLEN);
        char *row = myrow + cb * kTileBytes;
        if (valid == TLEN) {   // whole tile: registers in, registers out, no per-sample predicates
            V v[kVecPerRow];
#pragma unroll
            for (int i = 0; i < kVecPerRow; ++i) v[i] = *reinterpret_cast<const V *>(row + ((i ^ key) << 4));
#pragma unroll
            for (int i = 0; i < kVecPerRow; ++i) {
                T *e = reinterpret_cast<T *>(&v[i]);
#pragma unroll
                for (int j = 0; j < EPV; ++j) e[j] = (T)tf_step<NZ>((double)e[j], b, a, z);
            }
#pragma unroll
            for (int i = 0; i < kVecPerRow; ++i) *reinterpret_cast<V *>(row + ((i ^ key) << 4)) = v[i];
        } else {
            for (int i = 0; i < valid; ++i) {
                const int byte = i * (int)sizeof(T);
                T *e = reinterpret_cast<T *>(row + (((byte >> 4) ^ key) << 4) + (byte & 15));
                *e = (T)tf_step<NZ>((double)*e, b, a, z);
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
        case 7: return (int)run_fos<7>(p, aligned, s);
        case 8: return (int)run_fos<8>(p, aligned, s);
        default: return (int)cudaErrorInvalidValue;
    }
}

template <typename T, int NZ>
static int run_tf(const TfParams &p, bool aligned, cudaStream_t s) {
    // two tiles in flight (deeper rings measured slower: 12.5 / 18.2 / 19.2 ms for 2 / 3 / 4 stages on 1024 rows x 240 k)
    const unsigned grid = (unsigned)((p.R + 31) / 32);
    if (aligned) tf_seq_kernel<T, NZ, true, 2><<<grid, 32, 2 * kTileBytes, s>>>(p);
    else tf_seq_kernel<T, NZ, false, 2><<<grid, 32, 2 * kTileBytes, s>>>(p);
    return (int)cudaGetLastError();
}

template <typename T>
static int launch_tf_t(const TfParams &p, bool aligned, cudaStream_t s) {
    switch (p.ntaps - 1) {
        case 0:   // pure gain: one (always zero) delay value keeps the state layout [R][max(taps - 1, 1)]
        case 1: return run_tf<T, 1>(p, aligned, s);
        case 2: return run_tf<T, 2>(p, aligned, s);
        case 3: return run_tf<T, 3>(p, aligned, s);
        case 4: return run_tf<T, 4>(p, aligned, s);
        case 5: return run_tf<T, 5>(p, aligned, s);
        case 6: return run_tf<T, 6>(p, aligned, s);
        case 7: return run_tf<T, 7>(p, aligned, s);
        default: return (int)cudaErrorInvalidValue;
    }
}

int launch_tf(const TfParams &p, int dtype, bool aligned, cudaStream_t s) {
    return dtype == PFB_F64 ? launch_tf_t<double>(p, aligned, s) : launch_tf_t<float>(p, aligned, s);
}

// ---------------------------------------------------------------------------------------------
// TMA path of the gammatone bank: launch planning and per-call tables (cheap: 2*order x 2*order matrices).
// ---------------------------------------------------------------------------------------------
namespace {

// one sample of the cascade on long double states (re, im per stage), returns nothing (states only)
void fos_step_ld(const double *c3, int order, long double *z, long double x) {
    const long double g = c3[0], cr = c3[1], ci = c3[2];
    long double yr = z[0] + g * x, yi = z[1];
    z[0] = cr * yr - ci * yi;
    z[1] = cr * yi + ci * yr;
    for (int s = 1; s < order; ++s) {
        yr = z[2 * s] + yr;
        yi = z[2 * s + 1] + yi;
        z[2 * s] = cr * yr - ci * yi;
        z[2 * s + 1] = cr * yi + ci * yr;
    }
}

// P = M^L of the zero-input state map, rounded to double
void fos_transition(const double *c3, int order, long long L, double *P) {
    const int S = 2 * order;
    std::vector<long double> M(S * S), R(S * S, 0.0L), T(S * S);
    for (int jc = 0; jc < S; ++jc) {
        long double z[16] = {};
        z[jc] = 1.0L;
        fos_step_ld(c3, order, z, 0.0L);
        for (int i = 0; i < S; ++i) M[i * S + jc] = z[i];
    }
    for (int i = 0; i < S; ++i) R[i * S + i] = 1.0L;
    auto matmul = [&](const std::vector<long double> &a, const std::vector<long double> &b, std::vector<long double> &c) {
        for (int i = 0; i < S; ++i)
            for (int jc = 0; jc < S; ++jc) {
                long double acc = 0;
                for (int k = 0; k < S; ++k) acc += a[i * S + k] * b[k * S + jc];
                c[i * S + jc] = acc;
            }
    };
    for (long long e = L; e > 0; e >>= 1) {
        if (e & 1) { matmul(R, M, T); R = T; }
        if (e > 1) { matmul(M, M, T); M = T; }
    }
    for (int i = 0; i < S * S; ++i) P[i] = (double)R[i];
}

// decay window: smallest W (multiple of 16) such that the state response to a unit sample W samples back is
// below tol relative to its peak (the poles all have |c| < 1; order repeated poles decay like t^(order-1) |c|^t)
long long fos_warm(const double *c3, int order, long long L, double tol) {
    long double z[16] = {};
    fos_step_ld(c3, order, z, 1.0L);
    long double peak = 0;
    for (long long m = 1; m < L; ++m) {
        long double e = 0;
        for (int i = 0; i < 2 * order; ++i) e += z[i] * z[i];
        if (e > peak) peak = e;
        if (e <= (long double)tol * tol * peak && m > 8 * order) return std::min<long long>(L, (m + 15) / 16 * 16);
        fos_step_ld(c3, order, z, 0.0L);
    }
    return L;
}

}  // namespace

// Per-bank tables of the chunked gammatone path, kept between calls: building them costs 4-30 ms of host time per
// call for 64-96 bands (long double matrix powers and decay windows), more than the kernels of a short block, and
// uploading them needed a stream synchronisation.  Keyed by device, order and the coefficient bytes; one device
// buffer [P | warm] per chunk length.
namespace {
struct FosTables {
    void *dev = nullptr;      // P [B][2 order][2 order] doubles, then warm [B] ints at offW
    size_t offW = 0;
};
struct FosPlan {
    std::vector<long long> wproxy;            // decay windows for the launch planner (independent of the chunk length)
    std::map<long long, FosTables> by_len;    // chunk length -> tables
};
std::mutex g_fos_mu;
std::map<std::string, FosPlan> g_fos_plans;
size_t g_fos_tables = 0;

void fos_flush_tables() {    // cudaFree waits for the kernels that still read the buffers
    for (auto &kv : g_fos_plans)
        for (auto &t : kv.second.by_len) cudaFree(t.second.dev);
    g_fos_plans.clear();
    g_fos_tables = 0;
}
}  // namespace

template <int ORDER, int NX>
static int run_fos_tma_nx(const FosTmaParams &p, const FosCoef &tab, const TmaParams &cp, const CUtensorMap &xmap,
                          const CUtensorMap &ymap, cudaStream_t s) {
    auto ka = fos_tma_kernel<ORDER, false, NX>;
    auto kb = fos_tma_kernel<ORDER, true, NX>;
    const int smem_a = kFosStages * NX * kTileBytes + 2 * kFosStages * 8;
    const int smem_b = kFosStages * NX * kTileBytes + p.nb * out_slot_bytes(2 * NX) + 2 * kFosStages * 8;
    cudaError_t e = cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b);
    if (e != cudaSuccess) return (int)e;
    const unsigned grid = (unsigned)p.CB * (unsigned)p.G * (unsigned)p.NW;
    const int threads = (p.nb + 1) * 32;
    if (p.J > 1) {
        ka<<<grid, threads, smem_a, s>>>(p, tab, xmap, ymap);
        constexpr int S = 2 * ORDER;
        constexpr int LPC = S <= 2 ? 2 : S <= 4 ? 4 : S <= 8 ? 8 : 16;
        sos_carry_kernel<double, ORDER><<<(unsigned)(((long long)p.Q * LPC + 127) / 128), 128, 0, s>>>(cp);
    }
    kb<<<grid, threads, smem_b, s>>>(p, tab, xmap, ymap);
    return (int)cudaGetLastError();
}
template <int ORDER>
static int run_fos_tma(const FosTmaParams &p, const FosCoef &tab, const TmaParams &cp, const CUtensorMap &xmap,
                       const CUtensorMap &ymap, cudaStream_t s, bool wide) {
    return wide ? run_fos_tma_nx<ORDER, 2>(p, tab, cp, xmap, ymap, s) : run_fos_tma_nx<ORDER, 1>(p, tab, cp, xmap, ymap, s);
}

// returns 0 on success, a cudaError_t otherwise; *n_done = samples filtered (0 when the shapes do not fit this
// path; a ragged end [n_done, n) is left to the caller, states carried)
int launch_fos_tma(const double *x, long long ldx, double *y, long long ldy, const double *coef_host, int B, int order,
                   double *states, long long n, int C, cudaStream_t s, long long *n_done) {
    *n_done = 0;
    const int S = 2 * order;
    if (B > (int)(kCoefBytes / 24) || n < 64 || getenv("PFB_NO_TMA")) return 0;
    int sm_count = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    // wide variant (512-byte output fragments, <= 11 band warps, one CTA per SM) for banks with enough bands
    bool wide = B >= 9 && order <= 6;
    if (const char *v = getenv("PFB_FOS_WIDE")) wide = atoi(v) != 0;
    const int nbmax = wide ? kFosWideBands : kFosMaxBands;
    const int tt = wide ? 32 : 16;
    const int G0 = (B + nbmax - 1) / nbmax;
    const int nb = (B + G0 - 1) / G0, G = (B + nb - 1) / nb;
    // chunks per signal: fill the machine (2 CTAs per SM), prefer few chunks (the pre-pass is cheap: decay windows
    // of gammatone filters are a few thousand samples)
    std::lock_guard<std::mutex> lock(g_fos_mu);
    std::string key(reinterpret_cast<const char *>(coef_host), (size_t)B * 24);
    key.append(reinterpret_cast<const char *>(&dev), sizeof(dev));
    key.append(reinterpret_cast<const char *>(&order), sizeof(order));
    // bounded: 64 banks / 256 chunk lengths, then start over
    if ((g_fos_plans.size() >= 64 && !g_fos_plans.count(key)) || g_fos_tables >= 256) fos_flush_tables();
    FosPlan &plan = g_fos_plans[key];
    if (plan.wproxy.empty()) {
        plan.wproxy.resize(B);
        for (int b = 0; b < B; ++b) plan.wproxy[b] = fos_warm(coef_host + 3 * b, order, 1 << 20, 1e-13);
    }
    const std::vector<long long> &wproxy = plan.wproxy;
    int J = 0;
    if (const char *v = getenv("PFB_FOS_J")) J = atoi(v);
    if (J <= 0) {
        const double slots = (wide ? 1.0 : 2.0) * sm_count;
        double best = 1e300;
        for (int cand = 0; cand < 13; ++cand) {
            const int j = cand < 5 ? (1 << cand) : 32 * (cand - 4);
            const int jr = std::min(j, 32), chn = 32 / jr;
            if (j > 1 && n / j / tt * tt < 64 * tt) break;
            const long long cb = (C + chn - 1) / chn;
            const double waves = (double)cb * G * (j / jr) / slots;
            double eff = waves >= 1 ? waves / std::ceil(waves) : waves;
            eff *= (double)C / (double)(cb * chn);
            double share = 0;
            if (j > 1)
                for (int b = 0; b < B; ++b) share += std::min((double)j * (double)wproxy[b], (double)n);
            share /= (double)B * (double)n;
            const double cost = (1.0 + 0.7 * share) / eff;
            if (cost < best) best = cost, J = j;
        }
    }
    const int JR = std::min(J, 32), CH = 32 / JR;
    if (JR * CH != 32 || J % JR) return (int)cudaErrorInvalidValue;
    const long long Lt = n / J / tt * tt;
    if (Lt < 4 * tt || 2 * Lt >= 0x7fffffffLL) return 0;
    const long long n_main = Lt * J, Q = (long long)C * B;
    CUtensorMap xmap, ymap;
    if (!make_x_map(&xmap, PFB_F64, x, Lt, J, C, ldx, JR, CH) ||
        !make_y_map(&ymap, PFB_F64, y, 2 * Lt, J, B, C, 2 * ldy, JR, CH))
        return 0;
    static thread_local FosCoef tab;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < 3; ++i) tab.c[b][i] = coef_host[3 * b + i];
    FosTmaParams p{};
    p.states = states;
    p.L = Lt;
    p.Q = (int)Q, p.B = B, p.C = C, p.J = J, p.JR = JR, p.CH = CH, p.NW = J / JR, p.CB = (C + CH - 1) / CH;
    p.nb = nb, p.G = G;
    p.y = y, p.ldy2 = 2 * ldy;
    TmaParams cp{};
    void *scratch = nullptr;
    cudaError_t e = cudaSuccess;
    if (J > 1) {
        FosTables &tb = plan.by_len[Lt];
        if (!tb.dev) {
            std::vector<double> P((size_t)B * S * S);
            std::vector<int> warm(B);
            for (int b = 0; b < B; ++b) {
                fos_transition(coef_host + 3 * b, order, Lt, &P[(size_t)b * S * S]);
                warm[b] = (int)std::min<long long>(fos_warm(coef_host + 3 * b, order, Lt, 1e-13), Lt);
            }
            const size_t P_bytes = P.size() * 8, w_bytes = (size_t)B * 4;
            tb.offW = (P_bytes + 255) / 256 * 256;
            e = cudaMalloc(&tb.dev, tb.offW + w_bytes);
            if (e == cudaSuccess) e = cudaMemcpy(tb.dev, P.data(), P_bytes, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy((char *)tb.dev + tb.offW, warm.data(), w_bytes, cudaMemcpyHostToDevice);
            // copies from pageable memory may return before the DMA has landed, and the kernels below run on streams
            // that do not synchronise with the NULL stream
            if (e == cudaSuccess) e = cudaStreamSynchronize(0);
            if (e != cudaSuccess) {
                if (tb.dev) cudaFree(tb.dev);
                plan.by_len.erase(Lt);
                return (int)e;
            }
            ++g_fos_tables;
        }
        const size_t zs_bytes = (size_t)Q * J * S * 8;
        e = cudaMallocAsync(&scratch, zs_bytes, s);
        if (e != cudaSuccess) return (int)e;
        p.zs = scratch;
        p.warm = (const int *)((char *)tb.dev + tb.offW);
        cp.zs = scratch;
        cp.P = tb.dev;
        cp.states = states;
        cp.Q = (int)Q, cp.B = B, cp.J = J, cp.Ktot = order, cp.k0 = 0;
    }
    int err;
    switch (order) {
        case 1: err = run_fos_tma<1>(p, tab, cp, xmap, ymap, s, wide); break;
        case 2: err = run_fos_tma<2>(p, tab, cp, xmap, ymap, s, wide); break;
        case 3: err = run_fos_tma<3>(p, tab, cp, xmap, ymap, s, wide); break;
        case 4: err = run_fos_tma<4>(p, tab, cp, xmap, ymap, s, wide); break;
        case 5: err = run_fos_tma<5>(p, tab, cp, xmap, ymap, s, wide); break;
        case 6: err = run_fos_tma<6>(p, tab, cp, xmap, ymap, s, wide); break;
        case 7: err = run_fos_tma<7>(p, tab, cp, xmap, ymap, s, wide); break;
        case 8: err = run_fos_tma<8>(p, tab, cp, xmap, ymap, s, wide); break;
        default: err = (int)cudaErrorInvalidValue;
    }
    if (scratch) cudaFreeAsync(scratch, s);
    if (err) return err;
    *n_done = n_main;
    return 0;
}

}  // namespace pfb

namespace pfb {

// GammatoneFilterbank.synthesize / delay (pyfilterbank/gammatone.py:323-340), fused: every band's real part is
// delayed by its own delay_samples[b] (the first samples come from the delay memory of the previous block), bands
// with delay 0 are multiplied by their complex phase factor instead, and the bands are summed in ascending order
// (the order of numpy's sum over axis 0).  One thread per output sample; the band reads are coalesced along time.
//   bands [C][B][ldb] complex128, out [C][ldo] complex128 (imaginary part 0 unless a band has delay 0),
//   mem [C][B][dmax] real in/out, par [B][4] = gain, delay, Re phase factor, Im phase factor
__global__ void __launch_bounds__(256) gt_synth_kernel(const double2 *bands, long long ldb, double2 *out, long long ldo,
                                                       const double *mem, const double *par, int B, int dmax, long long n,
                                                       int C) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (t >= n) return;
    double re = 0.0, im = 0.0;
    const double2 *bc = bands + (size_t)c * B * ldb;
    for (int b = 0; b < B; ++b) {
        const double g = par[4 * b];
        const long long d = (long long)par[4 * b + 1];
        if (d == 0) {
            const double v = __dmul_rn(bc[(size_t)b * ldb + t].x, g);
            re = __dadd_rn(re, __dmul_rn(v, par[4 * b + 2]));
            im = __dadd_rn(im, __dmul_rn(v, par[4 * b + 3]));
        } else {
            const double v = t < d ? mem[((size_t)c * B + b) * dmax + t] : __dmul_rn(bc[(size_t)b * ldb + t - d].x, g);
            re = __dadd_rn(re, v);
        }
    }
    out[(size_t)c * ldo + t] = make_double2(re, im);
}

// new delay memory: mem[c][b][i] = gain_b * Re band[c][b][n - d_b + i], i < d_b (after the synthesis kernel has read the old one)
__global__ void gt_delay_memory_kernel(const double2 *bands, long long ldb, double *mem, const double *par, int B, int dmax,
                                       long long n, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y, c = blockIdx.z;
    const long long d = (long long)par[4 * b + 1];
    if (i >= d || d > n) return;
    mem[((size_t)c * B + b) * dmax + i] = __dmul_rn(bands[((size_t)c * B + b) * ldb + n - d + i].x, par[4 * b]);
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
    if (order < 1 || order > 8)
        return pfb_set_error(PFB_ERR_INVALID, "pfb_fos_bank_c128: order must be 1..8 per call (longer cascades are chained by the caller)");
    if ((long long)C * B > 0x7fffffffLL) return pfb_set_error(PFB_ERR_INVALID, "pfb_fos_bank_c128: C*B too large");
    cudaStream_t s = (cudaStream_t)stream;
    // TMA path: 16-byte aligned input rows (output rows are complex = 16 bytes per element)
    if ((((uintptr_t)x | (uintptr_t)y) & 15) == 0 && (ldx * 8) % 16 == 0) {
        long long n_done = 0;
        const int err = pfb::launch_fos_tma(x, ldx, y, ldy, coef_host, B, order, states, n, C, s, &n_done);
        if (err) return pfb_set_error(PFB_ERR_CUDA, cudaGetErrorString((cudaError_t)err));
        if (n_done == n) return pfb_set_error(PFB_OK, "");
        x += n_done;          // ragged end: sequential kernel from the carried states
        y += 2 * n_done;
        n -= n_done;
    }
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
    int err = pfb::launch_tf(p, PFB_F64, aligned, (cudaStream_t)stream);
    if (err) return pfb_set_error(PFB_ERR_CUDA, cudaGetErrorString((cudaError_t)err));
    return pfb_set_error(PFB_OK, "");
}

int pfb_gt_synthesize_c128(const double *bands, int64_t ldb, double *out, int64_t ldo, double *delay_memory,
                           const double *par_host, int B, int dmax, int64_t n, int C, void *stream) {
    if (n == 0) return pfb_set_error(PFB_OK, "");
    if (!bands || !out || !delay_memory || !par_host || B <= 0 || C <= 0 || n < 0 || ldb < n || ldo < n || dmax < 1)
        return pfb_set_error(PFB_ERR_INVALID, "pfb_gt_synthesize_c128: bad argument");
    for (int b = 0; b < B; ++b)
        if (par_host[4 * b + 1] < 0 || par_host[4 * b + 1] > dmax || par_host[4 * b + 1] > (double)n)
            return pfb_set_error(PFB_ERR_INVALID, "pfb_gt_synthesize_c128: every delay must be <= dmax and <= n");
    cudaStream_t s = (cudaStream_t)stream;
    double *dpar = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&dpar, (size_t)B * 4 * sizeof(double), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpar, par_host, (size_t)B * 4 * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return pfb_set_error(PFB_ERR_CUDA, cudaGetErrorString(e));
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)C);
    pfb::gt_synth_kernel<<<grid, 256, 0, s>>>((const double2 *)bands, ldb, (double2 *)out, ldo, delay_memory, dpar, B, dmax, n, C);
    dim3 gridm((unsigned)((dmax + 127) / 128), (unsigned)B, (unsigned)C);
    pfb::gt_delay_memory_kernel<<<gridm, 128, 0, s>>>((const double2 *)bands, ldb, delay_memory, dpar, B, dmax, n, C);
    e = cudaGetLastError();
    cudaFreeAsync(dpar, s);
    if (e != cudaSuccess) return pfb_set_error(PFB_ERR_CUDA, cudaGetErrorString(e));
    return pfb_set_error(PFB_OK, "");
}

}  // extern "C"
