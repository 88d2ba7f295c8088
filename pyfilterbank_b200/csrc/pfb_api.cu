// pfb_api.cu -- host side of libpfb_sos.so: bank objects, launch planning, the C ABI of
// include/pfb_sos.h.  No CPU filtering path exists here: every entry point ends in a kernel launch
// or an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/pfb_sos.h"
#include "pfb_kernels.cuh"
#include "pfb_tma.cuh"
#include "pfb_tf.cuh"

namespace {

thread_local int g_status = PFB_OK;
thread_local std::string g_error;

int fail(int status, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_status = status;
    g_error = buf;
    return status;
}

#define PFB_CUDA(call)                                                                           \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? PFB_ERR_NO_DEVICE \
                                                                                     : PFB_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct Sweep {
    int k0, K;
};

struct Tables {          // device tables of one (arithmetic type, chunk length)
    std::vector<void *> P;    // per sweep: [rows][2K][2K]
    double *gains = nullptr;  // [rows][K][2] = (G_k, 1/G_k) of the gain-normalised arithmetic, or null
    std::vector<int *> warm;  // per sweep: [rows] or null
    long long warm_total = 0;
    double *H = nullptr;      // chunk-impulse-response tables of the GEMM-form pre-pass (pfb_tma.cuh), or null
    long long *Hoff = nullptr;
    size_t bytes = 0;         // device bytes held (cache accounting)
    mutable unsigned long long last_use = 0;   // cache clock of the last call that used the set (eviction order)
};

void free_tables(Tables &t) {
    for (void *p : t.P) cudaFree(p);
    for (int *p : t.warm) cudaFree(p);
    cudaFree(t.H);
    cudaFree(t.Hoff);
    t.P.clear();
    t.warm.clear();
    t.H = nullptr;
    t.Hoff = nullptr;
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Every device allocation of this translation unit goes through here, so callers (the streaming session's
// statistics, tests/test_gpu_stream.py) can verify that steady-state block processing allocates nothing.
std::atomic<long long> g_device_allocs{0};
cudaError_t counted_malloc(void **p, size_t bytes) {
    ++g_device_allocs;
    return cudaMalloc(p, bytes);
}
template <typename P> cudaError_t counted_malloc(P **p, size_t bytes) { return counted_malloc((void **)p, bytes); }
cudaError_t counted_malloc_async(void **p, size_t bytes, cudaStream_t s) {
    ++g_device_allocs;
    return cudaMallocAsync(p, bytes, s);
}

// Optional stages fused into one bank pass (TMA kernels only; see TmaParams).
struct FilterExtra {
    // SPL-weighting prefilter ahead of the bank: taps normalised (wa[0] == 1), wz [C][wnz] device delay values in/out
    int wnz = 0;
    double wb[8] = {}, wa[8] = {};
    double *wz = nullptr;
    // the LAST band of the bank stores every second sample (parity dec_phase) to y2 [C][ldy2]; its states are
    // dec_states [C][K][2], y and `states` hold the other B - 1 bands
    bool dec = false;
    int dec_phase = 0;
    void *y2 = nullptr;
    long long ldy2 = 0;
    void *dec_states = nullptr;
    long long n_done = 0;    // out: samples filtered (a multiple of the tile step; the caller finishes the ragged end)
};
constexpr int kNotApplicable = 1;   // bank_filter_locked with a FilterExtra: the shapes do not fit the fused path

}  // namespace

struct pfb_bank {
    int device = 0;
    int rows = 0, B = 0, K = 0, per_channel = 0, C_sos = 0;
    std::vector<double> sos;  // [rows][K][6]
    void *coef[2] = {nullptr, nullptr};   // [PFB_F32 / PFB_F64] device [rows][K][5]
    std::vector<double> coef_host;        // [rows][K][5] = b0 b1 b2 a1 a2
    std::map<void *, std::pair<void *, size_t>> workspace;   // per stream: chunk-state scratch of the TMA path
    std::vector<Sweep> sweeps;
    std::map<std::tuple<int, long long, int>, Tables> tables;
    std::mutex mu;
    pfb_launch_info last{};
    int sm_count = 148;
    std::vector<double> wproxy;   // per band: decay length used by the launch-shape cost model
    double warm_tol = 1e-11;   // relative L2 budget of the truncated pre-pass (0: always exact)
    // pfb_bank_timing_begin / _end: event pairs recorded around the kernels of the next `timing_left` calls
    struct TimeRec { cudaEvent_t a, b; int kind; };   // kind 0 pre-pass, 1 carry, 2 output pass, 3 whole call
    std::vector<TimeRec> trecs;
    std::vector<cudaEvent_t> tevents;
    int timing_left = 0, timing_calls = 0;
    // device staging buffers of pfb_bank_filter_host (two input and two output slots), kept between calls
    void *hx[2] = {nullptr, nullptr}, *hy[2] = {nullptr, nullptr};
    size_t hx_bytes = 0, hy_bytes = 0;
    // PFB_ARITH_AUTO: per-band choice of arithmetic for float32 signals (calibrated once, ensure_auto_arith)
    int auto_state = 0;                 // 0 not calibrated, 1 ready, -1 failed
    int f32_bands = 0;
    unsigned *f32mask_dev = nullptr;    // bit b: band b runs in float32 arithmetic
    float *cf32_dev = nullptr;          // [B][K][5] plain coefficients as float
    int ensure_auto_arith();
    bool auto_now = false;              // the call in progress asked for PFB_ARITH_AUTO (set under `mu`)
    unsigned long long table_clock = 0; // ensure_tables: use counter of the table cache
    double slot_share = 1.0;            // share of the machine the chunk-count model may assume (stages that run side by side)
    int force_J = 0;                    // chunk count imposed by the owner of the bank (0: cost model)
    double *gains_dev[2] = {nullptr, nullptr};   // gain tables when no chunk tables are needed (single chunk)
    double *gains_only(int arith);
    // streams and events of the host-buffer entry points (pfb_bank_filter_host, pfb_bank_levels_host), created once
    // per bank: the per-stream scratch above is keyed by stream handle, so these must not change between calls
    std::mutex host_mu;                          // one host-buffer call per bank at a time (they share the staging slots)
    cudaStream_t hs[3] = {nullptr, nullptr, nullptr};   // compute, drain (D2H), input (H2D)
    cudaEvent_t hev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int ensure_host_ctx();

    cudaEvent_t tev_new() {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        tevents.push_back(e);
        return e;
    }
};

int pfb_bank::ensure_host_ctx() {
    if (hs[0]) return PFB_OK;
    for (cudaStream_t &st : hs) PFB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (cudaEvent_t &e : hev) PFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return PFB_OK;
}

// Gain tables of the normalised float64 arithmetic (single-sweep banks with well-scaled b0 only):
// [rows][K][2] = (G_k, 1 / G_k), G_k = b0_0 ... b0_{k-1}.  Null: plain arithmetic.
// PFB_ARITH_AUTO calibration: every band's cascade on 32768 samples of pseudo-random noise, once in the float32 arithmetic
// of the kernels (Biquad<float, false>::step: fused multiply-adds, float coefficients and states) and once in float64;
// a band runs in float32 where the two agree to 1e-5 relative L2 -- a factor 10 under the 1e-4 the float32 path has to
// meet against the double result.  Rounding noise of a stable section is stationary, so the ratio does not grow with
// the signal length; the low bands fail it by orders of magnitude (poles at radius 0.9999, angle 1e-3).
int pfb_bank::ensure_auto_arith() {
    if (auto_state != 0) return auto_state > 0 ? PFB_OK : fail(PFB_ERR_CUDA, "PFB_ARITH_AUTO: calibration tables unavailable");
    if (per_channel) {
        auto_state = -1;
        return fail(PFB_ERR_INVALID, "PFB_ARITH_AUTO needs a bank with shared coefficients");
    }
    const int n = 32768;
    std::vector<double> x(n);
    unsigned long long lcg = 0x9E3779B97F4A7C15ULL;
    for (int i = 0; i < n; ++i) {   // sum of four uniforms: close enough to Gaussian noise for a rounding-error estimate
        double v = 0;
        for (int r = 0; r < 4; ++r) {
            lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
            v += (double)(lcg >> 11) / 9007199254740992.0 - 0.5;
        }
        x[i] = 0.1 * v;
    }
    std::vector<unsigned> mask((B + 31) / 32, 0u);
    std::vector<float> cf((size_t)B * K * 5);
    f32_bands = 0;
    // acceptance threshold of the calibration: 1e-5, a factor 10 under the float32 bar (PFB_AUTO_TOL overrides it for
    // experiments: 3e-5 puts 14 instead of 12 bands of the 48 kHz third-octave bank on float32, profiles/r04y_auto_tol.txt)
    double accept = 1e-5;
    if (const char *v = getenv("PFB_AUTO_TOL")) {
        const double t = atof(v);
        if (t > 0 && t <= 1e-4) accept = t;
    }
    for (int b = 0; b < B; ++b) {
        const double *c = &coef_host[(size_t)b * K * 5];
        for (int i = 0; i < K * 5; ++i) cf[(size_t)b * K * 5 + i] = (float)c[i];
        std::vector<double> w64(2 * K, 0.0);
        std::vector<float> w32(2 * K, 0.0f);
        double num = 0, den = 0;
        for (int i = 0; i < n; ++i) {
            double s64 = x[i];
            float s32 = (float)x[i];
            for (int k = 0; k < K; ++k) {
                const double *q = c + 5 * k;
                const double w0 = s64 - q[3] * w64[2 * k] - q[4] * w64[2 * k + 1];
                s64 = q[0] * w0 + q[1] * w64[2 * k] + q[2] * w64[2 * k + 1];
                w64[2 * k + 1] = w64[2 * k];
                w64[2 * k] = w0;
                const float *qf = &cf[((size_t)b * K + k) * 5];
                const float t = std::fmaf(-qf[4], w32[2 * k + 1], s32);
                const float u = std::fmaf(qf[1], w32[2 * k], qf[2] * w32[2 * k + 1]);
                const float w0f = std::fmaf(-qf[3], w32[2 * k], t);
                s32 = std::fmaf(qf[0], w0f, u);
                w32[2 * k + 1] = w32[2 * k];
                w32[2 * k] = w0f;
            }
            const double d = (double)s32 - s64;
            num += d * d;
            den += s64 * s64;
        }
        const bool ok = std::isfinite(num) && den > 0 && std::sqrt(num / den) <= accept;
        if (ok) {
            mask[b >> 5] |= 1u << (b & 31);
            ++f32_bands;
        }
    }
    cudaError_t e = counted_malloc(&f32mask_dev, mask.size() * sizeof(unsigned));
    if (e == cudaSuccess) e = counted_malloc(&cf32_dev, cf.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(f32mask_dev, mask.data(), mask.size() * sizeof(unsigned), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(cf32_dev, cf.data(), cf.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);   // pageable copies may return before the DMA has landed
    if (e != cudaSuccess) {
        auto_state = -1;
        return fail(PFB_ERR_CUDA, "PFB_ARITH_AUTO: %s", cudaGetErrorString(e));
    }
    auto_state = 1;
    return PFB_OK;
}

double *pfb_bank::gains_only(int arith) {
    if (arith != PFB_F64 || sweeps.size() != 1 || getenv("PFB_NO_SCALED")) return nullptr;
    if (gains_dev[0]) return gains_dev[0];
    if (gains_dev[1]) return nullptr;   // marker: this bank's gains are out of range
    std::vector<double> g((size_t)rows * K * 2);
    bool ok = true;
    for (int r = 0; r < rows && ok; ++r) {
        double G = 1.0;
        for (int k = 0; k < K; ++k) {
            g[((size_t)r * K + k) * 2] = G;
            g[((size_t)r * K + k) * 2 + 1] = 1.0 / G;
            G *= coef_host[((size_t)r * K + k) * 5];
            const double m = std::fabs(G);
            if (!(m > 1e-200 && m < 1e200)) ok = false;
        }
    }
    if (!ok) {
        counted_malloc((void **)&gains_dev[1], 8);
        return nullptr;
    }
    if (counted_malloc((void **)&gains_dev[0], g.size() * sizeof(double)) != cudaSuccess) return nullptr;
    // (pageable-memory copies may return before the DMA has landed; the kernels run on streams that do not
    // synchronise with the NULL stream, hence the explicit wait -- once per bank)
    if (cudaMemcpy(gains_dev[0], g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaStreamSynchronize(0) != cudaSuccess) {
        cudaFree(gains_dev[0]);
        gains_dev[0] = nullptr;
        return nullptr;
    }
    return gains_dev[0];
}

namespace {

// Chunk transition of sections [k0, k0+Ks) of one coefficient row: P = M^L where M maps the
// state (w1_0, w2_0, w1_1, ...) one sample forward with zero input.  Built in long double
// (64-bit mantissa) by repeated squaring and rounded once to the arithmetic type.
void chunk_transition(const double *sos_row, int k0, int Ks, long long L, long double *R) {
    const int S = 2 * Ks;
    std::vector<long double> M(S * S), T(S * S);
    for (int j = 0; j < S; ++j) {
        long double st[32];
        for (int i = 0; i < S; ++i) st[i] = (i == j);
        long double x = 0;
        for (int k = 0; k < Ks; ++k) {
            const double *q = sos_row + 6 * (k0 + k);
            const long double w1 = st[2 * k], w2 = st[2 * k + 1];
            const long double w0 = x - (long double)q[4] * w1 - (long double)q[5] * w2;
            x = (long double)q[0] * w0 + (long double)q[1] * w1 + (long double)q[2] * w2;
            st[2 * k + 1] = w1;
            st[2 * k] = w0;
        }
        for (int i = 0; i < S; ++i) M[i * S + j] = st[i];
    }
    for (int i = 0; i < S * S; ++i) R[i] = 0;
    for (int i = 0; i < S; ++i) R[i * S + i] = 1;
    auto matmul = [&](const long double *a, const long double *b, long double *c) {
        for (int i = 0; i < S; ++i)
            for (int j = 0; j < S; ++j) {
                long double acc = 0;
                for (int k = 0; k < S; ++k) acc += a[i * S + k] * b[k * S + j];
                c[i * S + j] = acc;
            }
    };
    for (long long e = L; e > 0; e >>= 1) {
        if (e & 1) {
            matmul(R, M.data(), T.data());
            std::copy(T.begin(), T.end(), R);
        }
        if (e > 1) {
            matmul(M.data(), M.data(), T.data());
            M.swap(T);
        }
    }
}

// Pre-pass length of sections [k0, k0+Ks) for chunk length L.  A chunk hands its successor the
// zero-state response to its last W samples only; what is dropped is the response to the first
// L - W samples, i.e. the tail of the cascade's impulse response h beyond lag W.  For white input the
// error energy it leaves in the next chunk is  sum_{d=W+1..L} Tail(d),  Tail(d) = sum_{k>=d} h[k]^2,
// against a signal energy of  L * sum h^2.  Returns the smallest W (multiple of 32) whose relative
// L2 error estimate is <= tol, or L (exact pre-pass) if none is clearly shorter.
long long warm_length(const double *sos_row, int k0, int Ks, long long L, double tol) {
    if (tol <= 0 || L <= 64) return L;
    std::vector<double> tail((size_t)L + 2, 0.0);
    double w[16][2] = {};
    double total = 0, recent = 0;
    long long used = L + 1;
    for (long long n = 0; n <= L; ++n) {
        double x = n == 0 ? 1.0 : 0.0;
        for (int k = 0; k < Ks; ++k) {
            const double *q = sos_row + 6 * (k0 + k);
            const double w0 = x - q[4] * w[k][0] - q[5] * w[k][1];
            x = q[0] * w0 + q[1] * w[k][0] + q[2] * w[k][1];
            w[k][1] = w[k][0];
            w[k][0] = w0;
        }
        const double e = x * x;
        tail[n] = e;
        total += e;
        recent += e;
        if ((n & 1023) == 1023) {   // the response has died out: nothing later can matter
            if (recent <= 1e-60 * total) {
                used = n + 1;
                break;
            }
            recent = 0;
        }
    }
    if (!(total > 0) || !std::isfinite(total)) return L;
    // tail[d] := sum_{k >= d} h[k]^2 (within the simulated window; beyond it the response is < 1e-30 relative)
    for (long long d = used - 2; d >= 0; --d) tail[d] += tail[d + 1];
    // err2(W) = sum_{d=W+1..L} tail[d]; walk W downwards from L accumulating
    const double budget = tol * tol * (double)L * total;
    double acc = 0;
    long long W = L;
    for (long long d = std::min(L, used - 1); d >= 1; --d) {
        acc += tail[d];
        if (acc > budget) break;
        W = d - 1;
    }
    W = (W + 31) / 32 * 32;
    if (W < 32) W = 32;
    return W >= L - L / 8 ? L : W;
}

int ensure_coef(pfb_bank *b, int arith) {
    if (b->coef[arith]) return PFB_OK;
    const size_t n = (size_t)b->rows * b->K * 5;
    const std::vector<double> &c = b->coef_host;
    void *d = nullptr;
    if (arith == PFB_F64) {
        PFB_CUDA(counted_malloc(&d, n * sizeof(double)));
        PFB_CUDA(cudaMemcpy(d, c.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    } else {
        std::vector<float> f(c.begin(), c.end());
        PFB_CUDA(counted_malloc(&d, n * sizeof(float)));
        PFB_CUDA(cudaMemcpy(d, f.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    }
    PFB_CUDA(cudaStreamSynchronize(0));   // the upload has landed before any stream's kernel reads it
    b->coef[arith] = d;
    return PFB_OK;
}

// State impulse responses of one cascade: Hm[m * S + i] = state i (w1_0, w2_0, w1_1, ...) right after the
// sample that follows a unit input sample by m samples (m = 0: the sample itself), in long double.
void state_impulse_response(const double *sos_row, int Ks, long long W, std::vector<double> &Hm) {
    const int S = 2 * Ks;
    Hm.resize((size_t)W * S);
    long double w[16][2] = {};
    for (long long m = 0; m < W; ++m) {
        long double x = m == 0 ? 1.0L : 0.0L;
        for (int k = 0; k < Ks; ++k) {
            const double *q = sos_row + 6 * k;
            const long double w0 = x - (long double)q[4] * w[k][0] - (long double)q[5] * w[k][1];
            x = (long double)q[0] * w0 + (long double)q[1] * w[k][0] + (long double)q[2] * w[k][1];
            w[k][1] = w[k][0];
            w[k][0] = w0;
        }
        for (int k = 0; k < Ks; ++k) {
            Hm[(size_t)m * S + 2 * k] = (double)w[k][0];
            Hm[(size_t)m * S + 2 * k + 1] = (double)w[k][1];
        }
    }
}

// tt: samples per tile step of the TMA kernels (0: tables for the cp.async scan kernel, no GEMM-form tables)
int ensure_tables(pfb_bank *b, int arith, long long L, int tt, const Tables **out) {
    auto key = std::make_tuple(arith, L, tt);
    auto it = b->tables.find(key);
    if (it != b->tables.end()) {
        it->second.last_use = ++b->table_clock;
        *out = &it->second;
        return PFB_OK;
    }
    size_t cached_bytes = 0;
    for (auto &kv : b->tables) cached_bytes += kv.second.bytes;
    // Bound the cache (block-wise callers with a fixed block length hit it; the segment loop of the band-level path
    // and the host-streaming paths add a handful of chunk lengths each): over 96 sets or 8 GiB (PFB_TABLE_CACHE_MB) the
    // least recently used sets go until half the budget is free -- not all of them: a process that alternates between a
    // few shapes (bench.py: device-resident step, host-buffer step, band levels, float32 variants; up to 0.8 GB each)
    // would otherwise rebuild its tables on the host on every call (measured: 18 -> 320 ms per band-level call).
    // Nothing may still be using the evicted tables.
    const size_t cap = (size_t)std::max(64, env_int("PFB_TABLE_CACHE_MB", 8192)) << 20;
    if (b->tables.size() >= 96 || cached_bytes > cap) {
        PFB_CUDA(cudaDeviceSynchronize());
        while (!b->tables.empty() && (b->tables.size() >= 64 || cached_bytes > cap / 2)) {
            auto victim = b->tables.begin();
            for (auto jt = b->tables.begin(); jt != b->tables.end(); ++jt)
                if (jt->second.last_use < victim->second.last_use) victim = jt;
            cached_bytes -= std::min(cached_bytes, victim->second.bytes);
            free_tables(victim->second);
            b->tables.erase(victim);
        }
    }
    Tables t;
    t.last_use = ++b->table_clock;
    const auto t_build0 = std::chrono::steady_clock::now();
    // on a CUDA error the partial allocations held in t are released before returning
#define PFB_CUDA_T(call)                                                                         \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            free_tables(t);                                                                      \
            return fail(PFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                        \
    } while (0)
    // float32 arithmetic: 1e-6, two orders under its 1e-4 parity budget (float64: 1e-11 under 1e-9)
    const double tol = arith == PFB_F64 ? b->warm_tol : (b->warm_tol > 0 ? std::max(b->warm_tol, 1e-6) : 0.0);
    // rows with identical coefficients (a bank tiled per channel) share their tables
    std::map<std::string, int> first_row;
    std::vector<int> same_as(b->rows);
    for (int r = 0; r < b->rows; ++r) {
        std::string key((const char *)&b->sos[(size_t)r * b->K * 6], (size_t)b->K * 6 * sizeof(double));
        same_as[r] = first_row.emplace(std::move(key), r).first->second;
    }
    for (const Sweep &sw : b->sweeps) {
        const int S = 2 * sw.K;
        const size_t per = (size_t)S * S, n = per * b->rows;
        std::vector<long double> R(per);
        std::vector<double> Pd(n);
        std::vector<int> warm(b->rows);
        for (int r = 0; r < b->rows; ++r) {
            const int o = same_as[r];
            if (o != r) {
                std::copy(&Pd[o * per], &Pd[o * per] + per, &Pd[r * per]);
                warm[r] = warm[o];
            } else {
                const double *row = &b->sos[(size_t)r * b->K * 6];
                chunk_transition(row, sw.k0, sw.K, L, R.data());
                for (size_t i = 0; i < per; ++i) Pd[r * per + i] = (double)R[i];
                warm[r] = (int)std::min<long long>(warm_length(row, sw.k0, sw.K, L, tol), 0x7fffffff);
            }
            t.warm_total += std::min<long long>(warm[r], L);
        }
        int *dw = nullptr;
        PFB_CUDA_T(counted_malloc(&dw, b->rows * sizeof(int)));
        PFB_CUDA_T(cudaMemcpy(dw, warm.data(), b->rows * sizeof(int), cudaMemcpyHostToDevice));
        t.warm.push_back(dw);
        void *d = nullptr;
        if (arith == PFB_F64) {
            PFB_CUDA_T(counted_malloc(&d, n * sizeof(double)));
            PFB_CUDA_T(cudaMemcpy(d, Pd.data(), n * sizeof(double), cudaMemcpyHostToDevice));
        } else {
            std::vector<float> Pf(Pd.begin(), Pd.end());
            PFB_CUDA_T(counted_malloc(&d, n * sizeof(float)));
            PFB_CUDA_T(cudaMemcpy(d, Pf.data(), n * sizeof(float), cudaMemcpyHostToDevice));
        }
        t.P.push_back(d);
    }
    t.gains = b->gains_only(arith);   // owned by the bank
    // GEMM-form pre-pass tables (always float64; shared bank, one sweep, at most 16 states)
    if (tt > 0 && b->sweeps.size() == 1 && !b->per_channel && 2 * b->K <= 16 &&
        !env_int("PFB_NO_HFORM", 0)) {
        const int S = 2 * b->K, MT = (S + 7) / 8;
        const long long nt = L / tt;
        std::vector<long long> off(b->rows);
        std::vector<int> hwarm(b->rows);
        PFB_CUDA_T(cudaMemcpy(hwarm.data(), t.warm[0], b->rows * sizeof(int), cudaMemcpyDeviceToHost));
        long long total = 0;
        for (int r = 0; r < b->rows; ++r) {
            off[r] = total;
            long long wt = ((long long)hwarm[r] + tt - 1) / tt;
            if (wt > nt) wt = nt;
            total += wt * tt * MT * 8;
        }
        // (cap: a 119-band order-6 bank on few channels x 30 s needs ~340 MB; beyond the cap the recurrence pre-pass runs)
        if (total > 0 && total * 8 <= ((long long)env_int("PFB_HFORM_MAX_MB", 1024) << 20)) {
            std::vector<double> Hh((size_t)total, 0.0), Hm;
            for (int r = 0; r < b->rows; ++r) {
                long long wt = ((long long)hwarm[r] + tt - 1) / tt;
                if (wt > nt) wt = nt;
                const long long W = wt * tt;                       // window [L - W, L), sample t <-> lag m = L - 1 - t
                state_impulse_response(&b->sos[(size_t)r * b->K * 6], b->K, W, Hm);
                double *dst = &Hh[(size_t)off[r]];
                for (long long tl = 0; tl < W; ++tl) {
                    const long long m = W - 1 - tl;
                    const long long step = tl >> 2;
                    const int k = (int)(tl & 3);
                    for (int i = 0; i < S; ++i)
                        dst[(size_t)(step * MT + i / 8) * 32 + (i % 8) * 4 + k] = Hm[(size_t)m * S + i];
                }
            }
            t.bytes += (size_t)total * sizeof(double);
            PFB_CUDA_T(counted_malloc((void **)&t.H, (size_t)total * sizeof(double)));
            PFB_CUDA_T(cudaMemcpy(t.H, Hh.data(), (size_t)total * sizeof(double), cudaMemcpyHostToDevice));
            PFB_CUDA_T(counted_malloc((void **)&t.Hoff, b->rows * sizeof(long long)));
            PFB_CUDA_T(cudaMemcpy(t.Hoff, off.data(), b->rows * sizeof(long long), cudaMemcpyHostToDevice));
        }
    }
    PFB_CUDA_T(cudaStreamSynchronize(0));   // pageable uploads have landed before any stream's kernel reads the tables
#undef PFB_CUDA_T
    if (env_int("PFB_TABLE_TRACE", 0))   // which shapes build table sets, how large they are and how long the host takes
        fprintf(stderr, "[pfb tables] built set (arith %d, chunk %lld, tile step %d): %.1f MB in %.1f ms; cache now %zu sets, %.1f MB\n",
                arith, L, tt, t.bytes / 1048576.0,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build0).count(),
                b->tables.size() + 1, (cached_bytes + t.bytes) / 1048576.0);
    auto ins = b->tables.emplace(key, std::move(t));
    *out = &ins.first->second;
    return PFB_OK;
}

std::vector<Sweep> split_sweeps(int K) {
    static const int sizes[] = {8, 6, 4, 3, 2, 1};
    std::vector<Sweep> s;
    int k0 = 0;
    while (k0 < K) {
        for (int sz : sizes)
            if (sz <= K - k0) {
                s.push_back({k0, sz});
                k0 += sz;
                break;
            }
    }
    return s;
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

}  // namespace

// Chunked views of the signals for the TMA kernels.  A cascade's n_main = L * J samples are J chunks of L; a tile
// holds JR chunks of CH channels.  Boxes are one 128-byte half row wide with the 128-byte swizzle (16-byte chunk c
// of tile row r at c ^ (r & 7), the layout filter_half reads).
//   x [C][ldx]      -> 3-D (time L, chunk J, channel C),          box (tlen, JR, CH)
//   y [C][B][ldy]   -> 4-D (time L, chunk J, band B, channel C),  box (tlen, JR, 1, CH)
bool pfb::make_x_map(CUtensorMap *map, int dtype, const void *base, long long L, int J, long long C, long long ld, int JR, int CH) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t esz = dtype == PFB_F64 ? 8 : 4;
    cuuint64_t dims[3] = {(cuuint64_t)L, (cuuint64_t)J, (cuuint64_t)C};
    cuuint64_t strides[2] = {(cuuint64_t)L * esz, (cuuint64_t)ld * esz};
    cuuint32_t box[3] = {(cuuint32_t)(pfb::kRowBytes / esz), (cuuint32_t)JR, (cuuint32_t)CH};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, dtype == PFB_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                     const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
bool pfb::make_y_map(CUtensorMap *map, int dtype, void *base, long long L, int J, int B, long long C, long long ld, int JR, int CH) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t esz = dtype == PFB_F64 ? 8 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)L, (cuuint64_t)J, (cuuint64_t)B, (cuuint64_t)C};
    cuuint64_t strides[3] = {(cuuint64_t)L * esz, (cuuint64_t)ld * esz, (cuuint64_t)B * (cuuint64_t)ld * esz};
    cuuint32_t box[4] = {(cuuint32_t)(pfb::kRowBytes / esz), (cuuint32_t)JR, 1, (cuuint32_t)CH};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, dtype == PFB_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base,
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// TMA path helpers (pfb_tma.cuh): band grouping, tile-shape cost model, one launch set.
// ------------------------------------------------------------------------------------------------
static void tma_bands(const pfb_bank *b, int *nb, int *G, bool wide = false) {
    int nbmax = env_int("PFB_TMA_NB", wide ? pfb::kTmaMaxBands : pfb::kTmaSmallBands);
    nbmax = std::max(1, std::min(nbmax, pfb::kTmaMaxBands));
    const int G0 = (b->B + nbmax - 1) / nbmax;
    *nb = (b->B + G0 - 1) / G0;
    *G = (b->B + *nb - 1) / *nb;
}

// Tile shape: J chunks per cascade, a tile's 32 rows = JR chunks x CH channels.  Candidates: J = 1, 2, .. 16
// (JR = J, CH = 32 / J) and J = 32 g (JR = 32, CH = 1), at most jmax.  Cost model: output pass / wave efficiency
// of its grid, plus the pre-pass share estimated from the bands' decay lengths.
static int tma_pick_J(pfb_bank *b, long long n, int C, int tt, bool arith64, int G, long long jmax, bool wide = false) {
    int J = env_int("PFB_TMA_J", 0);
    if (env_int("PFB_TMA_GROUPS", 0) > 0) J = 32 * env_int("PFB_TMA_GROUPS", 0);
    if (J > 0 && J <= jmax) return J;
    if (b->force_J > 0 && b->force_J <= jmax) return b->force_J;
    J = 1;
    if (b->wproxy.empty()) {   // decay length of every band (pre-pass window if chunks were unbounded)
        b->wproxy.resize(b->B);
        for (int r = 0; r < b->B; ++r)
            b->wproxy[r] = (double)warm_length(&b->sos[(size_t)r * b->K * 6], 0, b->K, 1LL << 20, arith64 ? 1e-11 : 1e-6);
    }
    const double slots = (wide ? 1.0 : 2.0) * b->sm_count * b->slot_share;   // resident output-pass CTAs
    double best = 1e300;
    for (int cand = 0; cand < 13; ++cand) {
        const int j = cand < 5 ? (1 << cand) : 32 * (cand - 4);
        const int jr = std::min(j, 32), chn = 32 / jr;
        if (j > jmax) break;
        if (n / j / tt * tt < (j > 1 ? 64LL : 4LL) * tt && j > 1) break;
        const long long cb = (C + chn - 1) / chn;
        const double ctas = (double)cb * G * (j / jr);
        const double waves = ctas / slots;
        double eff = waves >= 1 ? waves / std::ceil(waves) : waves;
        eff *= (double)C / (double)(cb * chn);                         // ragged channel blocks idle lanes
        double share = 0;
        if (j > 1)
            for (int r = 0; r < b->B; ++r) share += std::min((double)j * b->wproxy[r], (double)n);
        share /= (double)b->B * (double)n;
        const double cost = (1.0 + 0.4 * share) / eff;
        if (cost < best) best = cost, J = j;
    }
    return J;
}

// One (pre-pass, carry, output | levels) launch set over J chunks of Lt samples.  Caller holds b->mu.
static int tma_run(pfb_bank *b, int dtype, bool arith64, const void *x, long long ldx, void *y, long long ldy, void *states,
                   long long Lt, int J, int C, int nb, int G, cudaStream_t stream, double *levels, long long lev_ld,
                   long long lev_off, int ut, bool may_time, bool wide = false, const FilterExtra *ex = nullptr, int jr_req = 0) {
    const int arith = arith64 ? PFB_F64 : PFB_F32;
    const int tt = 2 * pfb::kRowBytes / (dtype == PFB_F64 ? 8 : 4);
    // rows of a tile: JR chunks x CH channels; default JR = min(J, 32), a smaller JR (J = JR * chunk groups) spreads a
    // tile over more channels -- the chunked regime then has CTA counts other than multiples of 32 * groups
    const int JR = jr_req > 0 ? jr_req : std::min(J, 32), CH = 32 / JR;
    const int Q = C * b->B;
    const Tables *tab = nullptr;
    if (J > 1) {
        int rc = ensure_tables(b, arith, Lt, tt, &tab);
        if (rc) return rc;
    }
    const size_t asz = arith64 ? 8 : 4;
    // chunk-state scratch: kept per (bank, stream) so repeated calls do not allocate
    const size_t need = (size_t)Q * J * 2 * b->K * asz;
    auto &ws = b->workspace[(void *)stream];
    if (ws.second < need) {
        if (ws.first) {
            PFB_CUDA(cudaStreamSynchronize(stream));
            cudaFree(ws.first);
            ws = {nullptr, 0};
        }
        PFB_CUDA(counted_malloc(&ws.first, need));
        ws.second = need;
    }
    const bool timed = may_time && b->timing_left > 0;
    if (timed) { --b->timing_left; ++b->timing_calls; }
    const bool dec = ex && ex->dec;
    const bool weighted = ex && ex->wnz > 0;
    CUtensorMap xmap, ymap, ymap2;
    if (!pfb::make_x_map(&xmap, dtype, x, Lt, J, C, ldx, JR, CH)) return fail(PFB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    const int By = dec ? b->B - 1 : b->B;   // bands stored through the main output map
    if (y && By > 0) {
        if (!pfb::make_y_map(&ymap, dtype, y, Lt, J, By, C, ldy, JR, CH)) return fail(PFB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    } else {
        ymap = xmap;   // band-level epilogue / decimating band only: nothing is stored through it
    }
    ymap2 = ymap;
    if (dec && !pfb::make_y_map(&ymap2, dtype, ex->y2, Lt / 2, J, 1, C, ex->ldy2, JR, CH))
        return fail(PFB_ERR_CUDA, "cuTensorMapEncodeTiled failed (decimated output)");
    double *wz_copy = nullptr;
    if (weighted) {   // the kernel reads the start values from a copy: band group 0 overwrites wz while others may not have started
        PFB_CUDA(counted_malloc_async((void **)&wz_copy, (size_t)C * 8 * sizeof(double), stream));
        PFB_CUDA(cudaMemcpy2DAsync(wz_copy, 8 * sizeof(double), ex->wz, ex->wnz * sizeof(double), ex->wnz * sizeof(double),
                                   (size_t)C, cudaMemcpyDeviceToDevice, stream));
    }
    pfb::TmaParams tp{};
    tp.zs = ws.first;
    tp.states = states;
    tp.y = y, tp.ldy = ldy, tp.By = By;
    if (b->auto_now && b->auto_state == 1 && dtype == PFB_F32 && arith64 && !weighted && !dec && !levels) {
        tp.f32mask = b->f32mask_dev;
        tp.cf32 = b->cf32_dev;
    }
    tp.L = Lt;
    tp.Q = Q, tp.B = b->B, tp.C = C, tp.J = J, tp.nb = nb, tp.G = G;
    tp.JR = JR, tp.CH = CH, tp.NW = J / JR, tp.CB = (C + CH - 1) / CH;
    tp.Ktot = b->K, tp.k0 = 0;
    tp.nb_a = nb;
    tp.levels = levels, tp.lev_ld = lev_ld, tp.lev_off = lev_off, tp.ut = ut;
    tp.wide = wide ? 1 : 0;
    if (weighted) {
        tp.weighted = 1;
        tp.wnz = ex->wnz;
        tp.wz_in = wz_copy;
        tp.wz = ex->wz;
        for (int k = 0; k < 8; ++k) tp.wb[k] = ex->wb[k], tp.wa[k] = ex->wa[k];
    }
    if (dec) {
        tp.dec_band1 = b->B;       // the last band
        tp.dec_phase = ex->dec_phase;
        tp.Bs = b->B - 1;
        tp.dec_states = ex->dec_states;
    }
    if (tab) {
        tp.P = tab->P[0];
        tp.warm = tab->warm[0];
        tp.gains = tab->gains;
        tp.H = tab->H;
        tp.Hoff = tab->Hoff;
        // the GEMM-form pre-pass balances best with 4 bands per CTA (sweep in profiles/)
        tp.nb_a = std::max(1, std::min(env_int("PFB_TMA_NB_A", tab->H ? 4 : nb), nb));
    } else {
        tp.gains = b->gains_only(arith);   // single chunk: no transition / pre-pass tables needed
    }
    tp.G_a = (b->B + tp.nb_a - 1) / tp.nb_a;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (timed)
        for (cudaEvent_t &e : ev) e = b->tev_new();
    int launches = 0;
    const double *ch = b->coef_host.data();
#if PFB_WARP_PROF
    static unsigned long long *prof_dev = nullptr;
    if (env_int("PFB_WARP_PROF", 0)) {
        if (!prof_dev) cudaMalloc(&prof_dev, 16 * 8 * 8);
        cudaMemsetAsync(prof_dev, 0, 16 * 8 * 8, stream);
        tp.prof = prof_dev;
    }
#endif
    int err = dtype == PFB_F64 ? pfb::launch_tma_dd(tp, b->K, ch, xmap, ymap, ymap2, stream, &launches, timed ? ev : nullptr, 3)
              : arith64       ? pfb::launch_tma_fd(tp, b->K, ch, xmap, ymap, ymap2, stream, &launches, timed ? ev : nullptr, 3)
                              : pfb::launch_tma_ff(tp, b->K, ch, xmap, ymap, ymap2, stream, &launches, timed ? ev : nullptr, 3);
    if (wz_copy) cudaFreeAsync(wz_copy, stream);
    if (err != 0) return fail(PFB_ERR_CUDA, "TMA kernel launch failed: %s", cudaGetErrorString((cudaError_t)err));
#if PFB_WARP_PROF
    if (tp.prof) {   // experiment builds: clocks per warp role of the output pass, per CTA and tile step
        unsigned long long h[16 * 8];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, prof_dev, sizeof h, cudaMemcpyDeviceToHost);
        const double ctas = (double)tp.CB * G * tp.NW, steps = (double)(Lt / tt), d = ctas * steps;
        fprintf(stderr, "[warp prof] J=%d nb=%d G=%d ctas=%.0f tile steps=%.0f; clocks per tile step and CTA:\n"
                        "           wait-in    arith  out-wait    fence   arrive+store   total\n", J, nb, G, ctas, steps);
        for (int w = 0; w < 16; ++w)
            if (h[w * 8 + 3])
                fprintf(stderr, "  warp %2d %8.0f %8.0f %8.0f %8.0f %8.0f %12.0f%s\n", w, h[w * 8] / d, h[w * 8 + 1] / d, h[w * 8 + 2] / d,
                        h[w * 8 + 4] / d, h[w * 8 + 5] / d, h[w * 8 + 3] / d, w == 15 ? " (weighting)" : "");
    }
#endif
    if (timed) {
        b->trecs.push_back({ev[0], ev[1], 0});
        b->trecs.push_back({ev[1], ev[2], 1});
        b->trecs.push_back({ev[2], ev[3], 2});
        b->trecs.push_back({ev[0], ev[3], 3});
    }
    b->last = pfb_launch_info{};
    b->last.mode = PFB_MODE_SCAN;
    b->last.kernels = launches;
    b->last.warps_per_cascade = J / JR;
    b->last.chunk = Lt;
    b->last.warm_total = tab ? tab->warm_total * C : 0;
    // 2 = TMA path, 3 = with the wide (512-byte fragment) output pass, 4 = with the fused weighting warp,
    // 5 = with a decimating band
    b->last.aligned = weighted ? 4 : dec ? 5 : wide ? 3 : 2;
    return PFB_OK;
}

int pfb_set_error(int status, const char *msg) {   // used by the other translation units
    g_status = status;
    g_error = msg ? msg : "";
    return status;
}

extern "C" {

int pfb_last_status(void) { return g_status; }
const char *pfb_last_error(void) { return g_error.c_str(); }
const char *pfb_version(void) { return "pfb_sos 0.1 (sm_100a)"; }

int pfb_bank_create(pfb_bank **out, const double *sos, int C_sos, int B, int K, int per_channel) {
    if (!out || !sos || B <= 0 || K <= 0 || (per_channel && C_sos <= 0))
        return fail(PFB_ERR_INVALID, "pfb_bank_create: bad arguments (B=%d K=%d C_sos=%d)", B, K, C_sos);
    int dev = 0, ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PFB_ERR_NO_DEVICE, "no CUDA device (%s); libpfb_sos has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    PFB_CUDA(cudaGetDevice(&dev));
    pfb_bank *b = new pfb_bank;
    b->device = dev;
    b->B = B;
    b->K = K;
    b->per_channel = per_channel ? 1 : 0;
    b->C_sos = per_channel ? C_sos : 1;
    b->rows = b->C_sos * B;
    b->sos.assign(sos, sos + (size_t)b->rows * K * 6);
    b->sweeps = split_sweeps(K);
    b->coef_host.resize((size_t)b->rows * K * 5);
    for (size_t r = 0; r < (size_t)b->rows * K; ++r) {
        const double *q = &b->sos[r * 6];
        double *c = &b->coef_host[r * 5];
        c[0] = q[0], c[1] = q[1], c[2] = q[2];
        c[3] = q[4], c[4] = q[5];   // a0 = q[3] is ignored like the reference does (sosfilt.c:39)
    }
    cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (const char *v = getenv("PFB_WARM_TOL")) b->warm_tol = atof(v);
    *out = b;
    g_status = PFB_OK;
    return PFB_OK;
}

void pfb_bank_destroy(pfb_bank *b) {
    if (!b) return;
    for (int i = 0; i < 2; ++i)
        if (b->coef[i]) cudaFree(b->coef[i]);
    for (auto &kv : b->workspace) cudaFree(kv.second.first);
    for (cudaEvent_t e : b->tevents) cudaEventDestroy(e);
    cudaFree(b->gains_dev[0]);
    cudaFree(b->gains_dev[1]);
    cudaFree(b->f32mask_dev);
    cudaFree(b->cf32_dev);
    for (int i = 0; i < 2; ++i) {
        cudaFree(b->hx[i]);
        cudaFree(b->hy[i]);
    }
    for (auto &kv : b->tables) free_tables(kv.second);
    for (cudaStream_t st : b->hs)
        if (st) cudaStreamDestroy(st);
    for (cudaEvent_t e : b->hev)
        if (e) cudaEventDestroy(e);
    delete b;
}

int pfb_bank_release_stream(pfb_bank *b, void *stream) {
    if (!b) return fail(PFB_ERR_INVALID, "pfb_bank_release_stream: null bank");
    std::lock_guard<std::mutex> lock(b->mu);
    auto it = b->workspace.find(stream);
    if (it != b->workspace.end()) {
        if (it->second.first) {
            PFB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
            cudaFree(it->second.first);
        }
        b->workspace.erase(it);
    }
    g_status = PFB_OK;
    return PFB_OK;
}

long long pfb_device_allocs(void) { return g_device_allocs.load(); }

int pfb_bank_set_warm_tol(pfb_bank *b, double tol) {
    if (!b || !(tol >= 0)) return fail(PFB_ERR_INVALID, "pfb_bank_set_warm_tol: bad argument");
    std::lock_guard<std::mutex> lock(b->mu);
    if (tol != b->warm_tol) {
        cudaDeviceSynchronize();   // nothing may still be reading the tables built for the old tolerance
        for (auto &kv : b->tables) free_tables(kv.second);
        b->tables.clear();
        b->warm_tol = tol;
    }
    return PFB_OK;
}

int pfb_bank_timing_begin(pfb_bank *b, int max_calls) {
    if (!b || max_calls < 0) return fail(PFB_ERR_INVALID, "pfb_bank_timing_begin: bad argument");
    std::lock_guard<std::mutex> lock(b->mu);
    PFB_CUDA(cudaDeviceSynchronize());
    for (cudaEvent_t e : b->tevents) cudaEventDestroy(e);
    b->tevents.clear();
    b->trecs.clear();
    b->timing_left = max_calls;
    b->timing_calls = 0;
    return PFB_OK;
}

int pfb_bank_timing_end(pfb_bank *b, double *ms, int *calls) {
    if (!b || !ms || !calls) return fail(PFB_ERR_INVALID, "pfb_bank_timing_end: null argument");
    std::lock_guard<std::mutex> lock(b->mu);
    for (int i = 0; i < 4; ++i) ms[i] = 0;
    *calls = b->timing_calls;
    for (const pfb_bank::TimeRec &r : b->trecs) {
        PFB_CUDA(cudaEventSynchronize(r.b));
        float t;
        PFB_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.kind] += t;
    }
    if (b->timing_calls > 0)
        for (int i = 0; i < 4; ++i) ms[i] /= b->timing_calls;
    for (cudaEvent_t e : b->tevents) cudaEventDestroy(e);
    b->tevents.clear();
    b->trecs.clear();
    b->timing_left = b->timing_calls = 0;
    return PFB_OK;
}

static int check_device(const pfb_bank *b, const char *who);

int pfb_bank_f32_bands(pfb_bank *b) {
    if (!b) return fail(PFB_ERR_INVALID, "pfb_bank_f32_bands: null bank");
    std::lock_guard<std::mutex> lock(b->mu);
    if (int rc = check_device(b, "pfb_bank_f32_bands")) return rc;
    if (int rc = b->ensure_auto_arith()) return rc;
    return b->f32_bands;
}

int pfb_bank_last_launch(const pfb_bank *b, pfb_launch_info *info) {
    if (!b || !info) return fail(PFB_ERR_INVALID, "pfb_bank_last_launch: null argument");
    *info = b->last;
    return PFB_OK;
}

// pfb_bank_filter with the bank's mutex already held (the TMA path calls it again for the ragged end of a signal)
static int bank_filter_locked(pfb_bank *b, int dtype, const void *x, int64_t ldx, void *y, int64_t ldy, void *states,
                              int64_t n, int C, int flags, void *stream_, FilterExtra *ex = nullptr);

int pfb_bank_filter(pfb_bank *b, int dtype, const void *x, int64_t ldx, void *y, int64_t ldy, void *states,
                    int64_t n, int C, int flags, void *stream_) {
    if (!b) return fail(PFB_ERR_INVALID, "pfb_bank_filter: null bank");
    std::lock_guard<std::mutex> lock(b->mu);
    return bank_filter_locked(b, dtype, x, ldx, y, ldy, states, n, C, flags, stream_);
}

// A bank's coefficient, transition and scratch tables live on the device it was created on.
static int check_device(const pfb_bank *b, const char *who) {
    int dev = -1;
    PFB_CUDA(cudaGetDevice(&dev));
    if (dev != b->device)
        return fail(PFB_ERR_INVALID, "%s: the bank was created on CUDA device %d, the current device is %d", who, b->device, dev);
    return PFB_OK;
}

// ex != null: only the fused TMA path is tried -- returns kNotApplicable when the shapes do not fit it, otherwise
// filters ex->n_done samples (the caller finishes the ragged end).
static int bank_filter_locked(pfb_bank *b, int dtype, const void *x, int64_t ldx, void *y, int64_t ldy, void *states,
                              int64_t n, int C, int flags, void *stream_, FilterExtra *ex) {
    if (ex) ex->n_done = 0;
    if (n == 0) {   // nothing to filter: states stay as they are (sosfilt.c loads and stores them unchanged)
        g_status = PFB_OK;
        return PFB_OK;
    }
    if (!x || (!y && !(ex && ex->dec && b->B == 1)) || !states) return fail(PFB_ERR_INVALID, "pfb_bank_filter: null argument");
    if (int rc = check_device(b, "pfb_bank_filter")) return rc;
    if (dtype != PFB_F32 && dtype != PFB_F64) return fail(PFB_ERR_INVALID, "pfb_bank_filter: bad dtype %d", dtype);
    const bool dec = (flags & PFB_DECIMATE2) != 0;
    const int dec_phase = (flags & PFB_DEC_PHASE1) ? 1 : 0;
    const long long n_out = dec ? (n - dec_phase + 1) / 2 : n;
    if (n < 0 || C <= 0 || ldx < n || ldy < n_out)
        return fail(PFB_ERR_INVALID, "pfb_bank_filter: bad sizes n=%lld C=%d ldx=%lld ldy=%lld", (long long)n, C,
                    (long long)ldx, (long long)ldy);
    if (dec && b->sweeps.size() != 1)
        return fail(PFB_ERR_INVALID, "pfb_bank_filter: PFB_DECIMATE2 needs K in {1,2,3,4,6,8}, bank has K=%d", b->K);
    if (b->per_channel && C != b->C_sos)
        return fail(PFB_ERR_INVALID, "pfb_bank_filter: bank holds per-channel coefficients for %d channels, got C=%d",
                    b->C_sos, C);
    const long long Qll = (long long)C * b->B;
    if (Qll > 0x7fffffffLL) return fail(PFB_ERR_INVALID, "pfb_bank_filter: C*B = %lld exceeds 2^31-1", Qll);
    const bool exact = (flags & PFB_EXACT) != 0;
    if (flags & PFB_ARITH_AUTO) flags |= PFB_ARITH_F64;       // per-band arithmetic lives on the float64-state path
    const bool arith64 = dtype == PFB_F64 || (flags & PFB_ARITH_F64);
    if (dtype == PFB_F64 && (flags & PFB_ARITH_F64)) flags &= ~(PFB_ARITH_F64 | PFB_ARITH_AUTO);
    b->auto_now = false;
    if ((flags & PFB_ARITH_AUTO) && !exact && !b->per_channel) {
        if (int rca = b->ensure_auto_arith()) return rca;
        b->auto_now = true;
    }
    if (exact && (flags & PFB_ARITH_F64))
        return fail(PFB_ERR_INVALID, "pfb_bank_filter: PFB_EXACT and PFB_ARITH_F64 are mutually exclusive");
    int mode = flags & PFB_MODE_MASK;
    if (mode != PFB_MODE_AUTO && mode != PFB_MODE_SEQ && mode != PFB_MODE_SCAN)
        return fail(PFB_ERR_INVALID, "pfb_bank_filter: bad mode %d", mode);
    if (exact && mode == PFB_MODE_SCAN)
        return fail(PFB_ERR_INVALID, "pfb_bank_filter: PFB_EXACT needs PFB_MODE_SEQ (the scan re-associates)");
    const bool in_place = (const void *)x == (const void *)y;
    if (in_place && (b->B != 1 || ldx != ldy || dec))
        return fail(PFB_ERR_INVALID, "pfb_bank_filter: x may alias y only for single-band banks with ldx == ldy");

    cudaStream_t stream = (cudaStream_t)stream_;
    const int esz = dtype == PFB_F64 ? 8 : 4;
    const int tlen = pfb::kRowBytes / esz;
    const int arith = arith64 ? PFB_F64 : PFB_F32;
    const int Q = (int)Qll;

    // TMA kernels (pfb_tma.cuh): shared-input banks, one sweep, 16-byte aligned rows.  They cover both regimes:
    // few cascades (many time chunks per cascade) and many cascades (tile rows spread over channels, down to a
    // single chunk per cascade = plain sequential filtering with warp-uniform coefficients and no pre-pass).
    const int tt = 2 * tlen;
    bool tma_ok = !(flags & 0x100) && !dec && !exact && !b->per_channel && b->sweeps.size() == 1 && !in_place &&
                  !env_int("PFB_NO_TMA", 0) && aligned16(x) && aligned16(y) && (ldx * esz) % 16 == 0 &&
                  (ldy * esz) % 16 == 0 && n >= 4LL * tt &&
                  b->B * b->K <= (arith64 ? pfb::CoefTable<double>::kMaxSections : pfb::CoefTable<float>::kMaxSections);
    if (ex) {
        if (ex->dec) tma_ok = tma_ok && dtype == PFB_F64 && aligned16(ex->y2) && (ex->ldy2 * esz) % 16 == 0;
        if (ex->wnz) tma_ok = tma_ok && ex->wnz >= 4 && ex->wnz <= 6;
        if (!tma_ok || (mode != PFB_MODE_AUTO && mode != PFB_MODE_SCAN)) return kNotApplicable;
        mode = PFB_MODE_SCAN;
    }
    if (mode == PFB_MODE_AUTO) {
        // enough independent cascades to fill the machine with one lane each -> no time split needed
        const long long lanes_to_fill = (long long)b->sm_count * 16 * 32;
        if (exact || n < 4 * tlen) mode = PFB_MODE_SEQ;
        else if (tma_ok) mode = PFB_MODE_SCAN;
        else mode = Qll >= lanes_to_fill / 2 ? PFB_MODE_SEQ : PFB_MODE_SCAN;
    }
    int rc = PFB_OK;
    if (mode == PFB_MODE_SEQ) {
        rc = ensure_coef(b, arith);
        if (rc) return rc;
    }

    int warps = 1;
    long long L = n;
    const Tables *tab = nullptr;
    if (mode == PFB_MODE_SCAN && tma_ok) {
        // Output-pass variant with 512-byte row fragments (sos_tma_wide_kernel: one CTA of up to 11 band warps per
        // SM, 16 KB output tiles).  HBM takes 512-byte fragments ~8 % faster than 256-byte ones
        // (profiles/r01_ubench_tma_store_rows.txt); measured on cfg 2: 12.45 -> 12.0 ms per step.  Used for float64
        // signals of few-channel banks (chunked regime) with enough bands to fill a CTA; PFB_TMA_WIDE=0/1 overrides.
        int nb, G;
        tma_bands(b, &nb, &G, false);
        int J = tma_pick_J(b, n, C, tt, arith64, G, 1 << 30, false);
        // (few channels: the wide kernel's one CTA per SM leaves SMs idle -- measured slower below ~24 channels,
        // scripts/wide_probe.py -- so it needs enough (channel, band group) pairs for two CTAs per SM at J <= 128)
        const int Gw = (b->B + pfb::kTmaMaxBands - 1) / pfb::kTmaMaxBands;
        // (float32 signals with float64 arithmetic too: 10.24 -> 9.31 ms per cfg 2 step, profiles/r03h_fd_wide.txt)
        bool wide = b->K <= 4 && b->B >= 9 && J >= 32 && (long long)C * Gw * 4 >= 2LL * b->sm_count;
        wide = env_int("PFB_TMA_WIDE", wide ? 1 : 0) != 0;
        if (ex) wide = false;               // the fused stages live in the 256-byte-fragment kernel
        // Per-band arithmetic (PFB_ARITH_AUTO): the 256-byte-fragment kernel with 11 bands per CTA, two CTAs per SM --
        // the float64 warps of two CTAs keep the FP64 pipe busier than one CTA's share of them does (cfg 2 output pass
        // 6.28 ms against 6.88 ms on the wide kernel and 7.39 ms with 7 bands per CTA, profiles/r04b_auto_arith_variants.txt)
        if (b->auto_now && dtype == PFB_F32 && arith64 && !ex && b->K <= 4) {
            wide = env_int("PFB_TMA_WIDE", 0) != 0;
            if (!wide) {
                tma_bands(b, &nb, &G, true);
                J = tma_pick_J(b, n, C, tt, arith64, G, 1 << 30, false);
            }
        }
        if (ex && ex->wnz) {
            J = 1;                          // the weighting recurrence is sequential in time: rows are whole channels
            // Bands per CTA (sos_tma_weighted_kernel: 9 band warps on three SM sub-partitions, the weighting warp alone
            // on the fourth, one CTA per SM): as many as fit, in equal groups -- every group repeats the prefilter, and
            // three band warps per sub-partition reach the FP64 issue limit.  PFB_TMA_NB overrides.
            {
                const int nbmax = std::max(1, std::min(env_int("PFB_TMA_NB", pfb::kTmaWeightBands), pfb::kTmaWeightBands));
                G = (b->B + nbmax - 1) / nbmax;
                nb = (b->B + G - 1) / G;
            }
        }
        const int ttb = wide ? 2 * tt : tt;
        int jr_plan = 0;
        if (wide) {
            tma_bands(b, &nb, &G, true);
            // Few chunks mean little pre-pass work; the wide kernel runs one CTA per SM, so the grid is sized in
            // CTAs per SM: two for float64 (bandwidth bound, late CTAs run faster, insensitive to the partial last
            // round), three and a half for float32 (issue bound: the partial round costs).  Sweeps in profiles/.
            if (env_int("PFB_TMA_J", 0) <= 0 && env_int("PFB_TMA_GROUPS", 0) <= 0) {
                const long long want = dtype == PFB_F64 ? 2LL * b->sm_count : 7LL * b->sm_count / 2;
                J = 32;
                while ((long long)C * G * (J / 32) < want && J < 256) J += 32;
                while (J > 32 && n / J / ttb * ttb < 64LL * ttb) J -= 32;
            } else {
                J = tma_pick_J(b, n, C, ttb, arith64, G, 1 << 30, true);
            }
        }
        int JR = std::min(J, 32);
        if (jr_plan > 0 && J % jr_plan == 0) JR = jr_plan;
        if (const int jr_env = env_int("PFB_TMA_JR", 0)) {
            if (jr_env > 0 && jr_env <= 32 && 32 % jr_env == 0 && J % jr_env == 0) JR = jr_env;
        }
        if (J <= 0 || J % JR != 0 || 32 % JR != 0) return fail(PFB_ERR_INVALID, "pfb_bank_filter: bad chunk count %d", J);
        const long long Lt = n / J / ttb * ttb;
        if (Lt >= 4LL * ttb && Lt < 0x7fffffffLL) {
            const long long n_main = Lt * J;
            rc = tma_run(b, dtype, arith64, x, ldx, y, ldy, states, Lt, J, C, nb, G, stream, nullptr, 0, 0, 0, true, wide, ex, JR);
            if (rc) return rc;
            if (ex) {                // fused stages: the caller finishes the ragged end
                ex->n_done = n_main;
                g_status = PFB_OK;
                return PFB_OK;
            }
            pfb_launch_info main_info = b->last;
            if (n_main < n) {        // the ragged end continues from the carried states
                rc = bank_filter_locked(b, dtype, (const char *)x + n_main * esz, ldx, (char *)y + n_main * esz, ldy,
                                        states, n - n_main, C, (flags & ~PFB_MODE_MASK) | 0x100 | PFB_MODE_AUTO, stream_);
                if (rc) return rc;
                main_info.kernels += b->last.kernels;
            }
            b->last = main_info;
            g_status = PFB_OK;
            return PFB_OK;
        }
    }
    if (ex) return kNotApplicable;
    if (mode == PFB_MODE_SCAN) {
        const long long want_warps = (long long)b->sm_count * 16;
        warps = env_int("PFB_SCAN_WARPS", 0);
        if (warps <= 0) {
            warps = 1;
            while (warps < 8 && (long long)Q * warps < want_warps) warps *= 2;
        }
        if (warps > 8) warps = 8;
        const long long chunks = 32LL * warps;
        L = (n + chunks - 1) / chunks;
        L = std::max<long long>(tlen, (L + tlen - 1) / tlen * tlen);
        rc = ensure_tables(b, arith, L, 0, &tab);
        if (rc) return rc;
    }

    pfb::SosParams p{};
    p.y = y;
    p.coef = b->coef[arith];
    p.states = states;
    p.n = n;
    p.ldy = ldy;
    p.L = L;
    p.Q = Q;
    p.B = b->B;
    p.per_channel = b->per_channel;
    p.Ktot = b->K;
    p.dec = dec ? 1 : 0;
    p.dec_phase = dec_phase;
    int launches = 0;
    cudaEvent_t t0 = nullptr;
    if (!(flags & 0x100) && b->timing_left > 0) {   // single-kernel paths: all time is booked on the output pass
        --b->timing_left;
        ++b->timing_calls;
        t0 = b->tev_new();
        cudaEventRecord(t0, stream);
    }
    for (size_t si = 0; si < b->sweeps.size(); ++si) {
        const Sweep &sw = b->sweeps[si];
        p.k0 = sw.k0;
        if (si == 0) {
            p.x = x;
            p.ldx = ldx;
            p.xdiv = b->B;   // all bands of a channel read the channel's row
        } else {             // later sections filter the band signal in place
            p.x = y;
            p.ldx = ldy;
            p.xdiv = 1;
        }
        p.aligned = aligned16(p.x) && aligned16(y) && (p.ldx * esz) % 16 == 0 && (ldy * esz) % 16 == 0;
        int err;
        if (mode == PFB_MODE_SEQ) {
            err = dtype == PFB_F64 ? pfb::launch_seq_dd(p, sw.K, exact, stream)
                  : arith64       ? pfb::launch_seq_fd(p, sw.K, false, stream)
                                  : pfb::launch_seq_ff(p, sw.K, exact, stream);
        } else {
            p.P = tab->P[si];
            p.warm = tab->warm[si];
            const double *ch = b->coef_host.data();
            err = dtype == PFB_F64 ? pfb::launch_scan_dd(p, sw.K, warps, ch, b->rows, C, stream, &launches)
                  : arith64       ? pfb::launch_scan_fd(p, sw.K, warps, ch, b->rows, C, stream, &launches)
                                  : pfb::launch_scan_ff(p, sw.K, warps, ch, b->rows, C, stream, &launches);
        }
        if (err != 0)
            return fail(PFB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString((cudaError_t)err));
        if (mode == PFB_MODE_SEQ) ++launches;
    }
    if (t0) {
        cudaEvent_t t1 = b->tev_new();
        cudaEventRecord(t1, stream);
        b->trecs.push_back({t0, t1, 2});
        b->trecs.push_back({t0, t1, 3});
    }
    b->last.mode = mode;
    b->last.kernels = launches;
    b->last.warps_per_cascade = warps;
    b->last.chunk = L;
    b->last.warm_total = tab ? tab->warm_total * C / b->C_sos : 0;
    b->last.aligned = p.aligned;
    g_status = PFB_OK;
    return PFB_OK;
}

// Band levels of HOST signals: channel groups stream through the device (input copy of group g+1 under the kernels
// of group g); only the energies travel back.  The path is bound by the host-to-device copy of the input.
int pfb_bank_levels_host(pfb_bank *b, int dtype, const void *x, int64_t ldx, double *levels, int64_t ld_levels,
                         void *states, int64_t n, int C, int64_t block_len, int flags) {
    if (!b || !x || !levels || !states) return fail(PFB_ERR_INVALID, "pfb_bank_levels_host: null argument");
    if (dtype != PFB_F32 && dtype != PFB_F64) return fail(PFB_ERR_INVALID, "pfb_bank_levels_host: bad dtype");
    if (n <= 0 || C <= 0 || ldx < n || block_len <= 0 || n % block_len != 0 || ld_levels < n / block_len)
        return fail(PFB_ERR_INVALID, "pfb_bank_levels_host: bad sizes");
    const size_t esz = dtype == PFB_F64 ? 8 : 4;
    if (flags & PFB_ARITH_AUTO) flags |= PFB_ARITH_F64;   // per-band arithmetic implies float64 states
    const size_t ssz = (dtype == PFB_F64 || (flags & PFB_ARITH_F64)) ? 8 : 4;
    const int B = b->B;
    const long long nblocks = n / block_len;
    const long long pitch = (n + 31) / 32 * 32;
    // groups of at most 1 GiB of input, at least four groups when there are that many channels
    long long Cg = std::max<long long>(1, std::min<long long>(C, (1LL << 30) / (long long)(pitch * esz)));
    Cg = std::min<long long>(Cg, std::max<long long>(1, (C + 3) / 4));
    void *dx[2] = {nullptr, nullptr}, *ds = nullptr;
    double *dl = nullptr;
    std::lock_guard<std::mutex> host_lock(b->host_mu);
    if (int rc0 = check_device(b, "pfb_bank_levels_host")) return rc0;
    if (int rc0 = b->ensure_host_ctx()) return rc0;
    cudaStream_t sc = b->hs[0], sh = b->hs[2];
    cudaEvent_t in_done[2] = {b->hev[0], b->hev[1]}, used[2] = {b->hev[2], b->hev[3]};
    int rc = PFB_OK;
    auto cleanup = [&]() {
        cudaStreamSynchronize(sc);
        cudaStreamSynchronize(sh);
        cudaFree(dx[0]); cudaFree(dx[1]); cudaFree(ds); cudaFree(dl);
    };
#define PFB_CUDA_L(call)                                                                                  \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            rc = fail(PFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            cleanup();                                                                                    \
            return rc;                                                                                    \
        }                                                                                                 \
    } while (0)
    for (int i = 0; i < 2; ++i) PFB_CUDA_L(counted_malloc(&dx[i], (size_t)Cg * pitch * esz));
    const size_t nstate = (size_t)C * B * b->K * 2;
    PFB_CUDA_L(counted_malloc(&ds, nstate * ssz));
    PFB_CUDA_L(counted_malloc((void **)&dl, (size_t)C * B * nblocks * sizeof(double)));
    PFB_CUDA_L(cudaMemcpyAsync(ds, states, nstate * ssz, cudaMemcpyHostToDevice, sc));
    const int ngroups = (int)((C + Cg - 1) / Cg);
    for (int g = 0; g < ngroups; ++g) {
        const int s = g & 1;
        const int c0 = (int)(g * Cg), cg = (int)std::min<long long>(Cg, C - c0);
        if (g >= 2) PFB_CUDA_L(cudaStreamWaitEvent(sh, used[s], 0));      // the slot's previous group has been filtered
        PFB_CUDA_L(cudaMemcpy2DAsync(dx[s], pitch * esz, (const char *)x + (size_t)c0 * ldx * esz, ldx * esz, n * esz, cg,
                                     cudaMemcpyHostToDevice, sh));
        PFB_CUDA_L(cudaEventRecord(in_done[s], sh));
        PFB_CUDA_L(cudaStreamWaitEvent(sc, in_done[s], 0));
        rc = pfb_bank_levels(b, dtype, dx[s], pitch, dl + (size_t)c0 * B * nblocks, nblocks,
                             (char *)ds + (size_t)c0 * B * b->K * 2 * ssz, n, cg, block_len, flags, sc);
        if (rc) {
            cleanup();
            return rc;
        }
        PFB_CUDA_L(cudaEventRecord(used[s], sc));
    }
    PFB_CUDA_L(cudaMemcpy2DAsync(levels, ld_levels * sizeof(double), dl, nblocks * sizeof(double), nblocks * sizeof(double),
                                 (size_t)C * B, cudaMemcpyDeviceToHost, sc));
    PFB_CUDA_L(cudaMemcpyAsync(states, ds, nstate * ssz, cudaMemcpyDeviceToHost, sc));
    PFB_CUDA_L(cudaStreamSynchronize(sc));
    PFB_CUDA_L(cudaStreamSynchronize(sh));
    cleanup();
    g_status = PFB_OK;
    return PFB_OK;
#undef PFB_CUDA_L
}

// Host-buffer filtering of a shared bank by CHANNEL GROUPS: a group's whole time range is filtered in one call, so
// its band signals [Cg][B][n] are one long-row 2-D (contiguous when ldy == n) device-to-host copy -- PCIe runs at
// the rate of plain contiguous copies (57 GB/s measured against 48 GB/s for time blocks, whose host side is
// 2112 scattered 0.5 MB pieces per block).  Input copy, kernels and output drain of consecutive groups overlap.
static int filter_host_by_channels(pfb_bank *b, int dtype, const void *x, int64_t ldx, void *y, int64_t ldy,
                                   void *states, int64_t n, int C, int flags, int Cg) {
    const size_t esz = dtype == PFB_F64 ? 8 : 4;
    if (flags & PFB_ARITH_AUTO) flags |= PFB_ARITH_F64;   // per-band arithmetic implies float64 states
    const size_t ssz = (dtype == PFB_F64 || (flags & PFB_ARITH_F64)) ? 8 : 4;
    const int B = b->B;
    const long long pitch = (n + 31) / 32 * 32;
    void *dx[2] = {nullptr, nullptr}, *dy[2] = {nullptr, nullptr}, *ds = nullptr;
    if (int rc0 = b->ensure_host_ctx()) return rc0;
    cudaStream_t sc = b->hs[0], sd = b->hs[1];
    cudaEvent_t done[2] = {b->hev[0], b->hev[1]}, drained[2] = {b->hev[2], b->hev[3]};
    int rc = PFB_OK;
    auto cleanup = [&]() {
        cudaStreamSynchronize(sc);
        cudaStreamSynchronize(sd);
        cudaFree(ds);
    };
#define PFB_CUDA_C(call)                                                                                  \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            rc = fail(PFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            cleanup();                                                                                    \
            return rc;                                                                                    \
        }                                                                                                 \
    } while (0)
    const int ngroups = (C + Cg - 1) / Cg;
    {   // staging buffers live in the bank: repeated calls (block-wise callers, the bench) do not reallocate
        std::lock_guard<std::mutex> lock(b->mu);
        const size_t need_x = (size_t)Cg * pitch * esz, need_y = (size_t)Cg * B * pitch * esz;
        if (b->hx_bytes < need_x || b->hy_bytes < need_y) {
            PFB_CUDA_C(cudaDeviceSynchronize());
            for (int i = 0; i < 2; ++i) {
                cudaFree(b->hx[i]);
                cudaFree(b->hy[i]);
                b->hx[i] = b->hy[i] = nullptr;
            }
            b->hx_bytes = b->hy_bytes = 0;
            for (int i = 0; i < 2; ++i) {
                PFB_CUDA_C(counted_malloc(&b->hx[i], need_x));
                PFB_CUDA_C(counted_malloc(&b->hy[i], need_y));
            }
            b->hx_bytes = need_x;
            b->hy_bytes = need_y;
        }
        for (int i = 0; i < 2; ++i) {
            dx[i] = b->hx[i];
            dy[i] = b->hy[i];
        }
    }
    // all states stay on the device for the whole call (the host array may be pageable: one copy each way)
    const size_t nstate = (size_t)C * B * b->K * 2;
    PFB_CUDA_C(counted_malloc(&ds, nstate * ssz));
    PFB_CUDA_C(cudaMemcpyAsync(ds, states, nstate * ssz, cudaMemcpyHostToDevice, sc));
    for (int g = 0; g < ngroups; ++g) {
        const int s = g & 1;
        const int c0 = g * Cg, cg = std::min(Cg, C - c0);
        char *dsg = (char *)ds + (size_t)c0 * B * b->K * 2 * ssz;
        if (g >= 2) PFB_CUDA_C(cudaStreamWaitEvent(sc, drained[s], 0));   // the slot's previous output is out
        PFB_CUDA_C(cudaMemcpy2DAsync(dx[s], pitch * esz, (const char *)x + (size_t)c0 * ldx * esz, ldx * esz, n * esz, cg,
                                     cudaMemcpyHostToDevice, sc));
        rc = pfb_bank_filter(b, dtype, dx[s], pitch, dy[s], pitch, dsg, n, cg, flags, sc);
        if (rc) {
            cleanup();
            return rc;
        }
        PFB_CUDA_C(cudaEventRecord(done[s], sc));
        PFB_CUDA_C(cudaStreamWaitEvent(sd, done[s], 0));
        PFB_CUDA_C(cudaMemcpy2DAsync((char *)y + (size_t)c0 * B * ldy * esz, ldy * esz, dy[s], pitch * esz, n * esz,
                                     (size_t)cg * B, cudaMemcpyDeviceToHost, sd));
        PFB_CUDA_C(cudaEventRecord(drained[s], sd));
    }
    PFB_CUDA_C(cudaMemcpyAsync(states, ds, nstate * ssz, cudaMemcpyDeviceToHost, sc));
    PFB_CUDA_C(cudaStreamSynchronize(sc));
    PFB_CUDA_C(cudaStreamSynchronize(sd));
    cleanup();
    g_status = PFB_OK;
    return PFB_OK;
#undef PFB_CUDA_C
}

static int levels_impl(pfb_bank *b, int dtype, const void *x, int64_t ldx, double *levels, int64_t ld_levels, void *states,
                       int64_t n, int C, int64_t block_len, int flags, void *stream_, FilterExtra *ex) {
    if (!b || !x || !levels || !states) return fail(PFB_ERR_INVALID, "pfb_bank_levels: null argument");
    if (dtype != PFB_F32 && dtype != PFB_F64) return fail(PFB_ERR_INVALID, "pfb_bank_levels: bad dtype %d", dtype);
    const int esz = dtype == PFB_F64 ? 8 : 4;
    const int tt = 2 * pfb::kRowBytes / esz;
    if (flags & PFB_ARITH_AUTO) flags |= PFB_ARITH_F64;
    const bool arith64 = dtype == PFB_F64 || (flags & PFB_ARITH_F64);
    if (n <= 0 || C <= 0 || ldx < n || block_len <= 0 || n % block_len != 0 || block_len % tt != 0 ||
        ld_levels < n / block_len)
        return fail(PFB_ERR_INVALID,
                    "pfb_bank_levels: n (%lld) must be a multiple of block_len (%lld), block_len a multiple of %d samples",
                    (long long)n, (long long)block_len, tt);
    if (b->per_channel || b->sweeps.size() != 1 || !aligned16(x) || (ldx * esz) % 16 != 0 ||
        b->B * b->K > (arith64 ? pfb::CoefTable<double>::kMaxSections : pfb::CoefTable<float>::kMaxSections))
        return fail(PFB_ERR_INVALID, "pfb_bank_levels: needs a shared bank with K in {1,2,3,4,6,8} and 16-byte aligned rows");
    if ((long long)C * b->B > 0x7fffffffLL) return fail(PFB_ERR_INVALID, "pfb_bank_levels: C*B too large");
    std::lock_guard<std::mutex> lock(b->mu);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long nblocks = n / block_len, Q = (long long)C * b->B;
    // energy unit: block_len / r, r the smallest divisor of block_len / tt that gives enough units to cut into chunks
    long long r = 1;
    const long long bt = block_len / tt;
    for (long long d = 1; d <= bt; ++d) {
        if (bt % d) continue;
        r = d;
        if (nblocks * d >= 512 || block_len / d < 8LL * tt) break;
    }
    const long long unit = block_len / r, nunits = nblocks * r;
    double *partial = levels;
    long long pld = ld_levels;
    if (r > 1) {
        PFB_CUDA(counted_malloc_async((void **)&partial, (size_t)Q * nunits * sizeof(double), stream));
        pld = nunits;
    }
    int nb, G;
    tma_bands(b, &nb, &G);
    if (ex && ex->wnz) {   // fused weighting prefilter: rows are whole channels, <= 9 bands per CTA (sos_tma_weighted_kernel)
        G = (b->B + pfb::kTmaWeightBands - 1) / pfb::kTmaWeightBands;
        nb = (b->B + G - 1) / G;
    }
    const long long min_units = (4LL * tt + unit - 1) / unit;   // a chunk is at least four tile steps
    long long pos = 0;
    int kernels = 0;
    while (pos < nunits) {
        const long long rem = nunits - pos;
        int J = ex && ex->wnz ? 1 : tma_pick_J(b, rem * unit, C, tt, arith64, G, std::max<long long>(1, rem / min_units));
        if (J > rem) J = 1;
        const int JR = std::min(J, 32);
        if (J % JR != 0 || 32 % JR != 0) J = 1;
        const long long Lu = rem / J;
        if (Lu * unit >= 0x7fffffffLL) return fail(PFB_ERR_INVALID, "pfb_bank_levels: signal too long for one call");
        int rc = tma_run(b, dtype, arith64, (const char *)x + pos * unit * esz, ldx, nullptr, 0, states, Lu * unit, J, C, nb, G,
                         stream, partial, pld, pos, (int)(unit / tt), pos == 0, false, ex);
        if (rc) return rc;
        kernels += b->last.kernels;
        pos += (long long)J * Lu;
    }
    if (r > 1) {
        const long long total = Q * nblocks;
        pfb::levels_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(partial, pld, levels, ld_levels, nblocks,
                                                                                       (int)r, Q);
        ++kernels;
        PFB_CUDA(cudaGetLastError());
        PFB_CUDA(cudaFreeAsync(partial, stream));
    }
    b->last.kernels = kernels;
    g_status = PFB_OK;
    return PFB_OK;
}

int pfb_bank_levels(pfb_bank *b, int dtype, const void *x, int64_t ldx, double *levels, int64_t ld_levels, void *states,
                    int64_t n, int C, int64_t block_len, int flags, void *stream_) {
    return levels_impl(b, dtype, x, ldx, levels, ld_levels, states, n, C, block_len, flags, stream_, nullptr);
}

// Weighted band levels: SPL-weighting prefilter (splweighting.py:61-80) + bank + energies per block in ONE kernel -- the
// sound-level-meter use of the reference's pieces (weight_signal, filter_mimo_c, 10 log10 sum y^2: octbank.py:262-263):
// neither the weighted signal nor a band signal is written, only the input is read.
int pfb_bank_levels_weighted(pfb_bank *b, int dtype, const void *x, int64_t ldx, double *levels, int64_t ld_levels,
                             void *states, const double *wb, int nwb, const double *wa, int nwa, double *wz, int64_t n, int C,
                             int64_t block_len, int flags, void *stream_) {
    if (!b || !wb || !wa || !wz || nwb < 1 || nwa < 1 || nwb > 8 || nwa > 8)
        return fail(PFB_ERR_INVALID, "pfb_bank_levels_weighted: bad argument (at most 8 taps)");
    if (wa[0] == 0.0) return fail(PFB_ERR_INVALID, "pfb_bank_levels_weighted: a[0] must not be 0");
    const int nz = std::max(std::max(nwb, nwa) - 1, 1);
    if (nz < 4 || nz > 6) return fail(PFB_ERR_INVALID, "pfb_bank_levels_weighted: the fused prefilter takes 5 to 7 taps");
    FilterExtra ex;
    for (int k = 0; k < 8; ++k) {   // scipy normalises by a[0] first
        ex.wb[k] = k < nwb ? wb[k] / wa[0] : 0.0;
        ex.wa[k] = k < nwa ? wa[k] / wa[0] : 0.0;
    }
    ex.wnz = nz;
    ex.wz = wz;
    return levels_impl(b, dtype, x, ldx, levels, ld_levels, states, n, C, block_len, flags, stream_, &ex);
}

int pfb_bank_filter_host(pfb_bank *b, int dtype, const void *x, int64_t ldx, void *y, int64_t ldy, void *states,
                         int64_t n, int C, int flags) {
    if (!b) return fail(PFB_ERR_INVALID, "pfb_bank_filter_host: null bank");
    if (n == 0) {
        g_status = PFB_OK;
        return PFB_OK;
    }
    if (!x || !y || !states) return fail(PFB_ERR_INVALID, "pfb_bank_filter_host: null argument");
    if (dtype != PFB_F32 && dtype != PFB_F64) return fail(PFB_ERR_INVALID, "pfb_bank_filter_host: bad dtype");
    if (n < 0 || C <= 0 || ldx < n || ldy < n) return fail(PFB_ERR_INVALID, "pfb_bank_filter_host: bad sizes");
    if (flags & PFB_DECIMATE2) return fail(PFB_ERR_INVALID, "pfb_bank_filter_host: PFB_DECIMATE2 is a device-pointer option");
    std::lock_guard<std::mutex> host_lock(b->host_mu);
    if (int rc0 = check_device(b, "pfb_bank_filter_host")) return rc0;
    const size_t esz = dtype == PFB_F64 ? 8 : 4;
    if (flags & PFB_ARITH_AUTO) flags |= PFB_ARITH_F64;   // per-band arithmetic implies float64 states
    const size_t ssz = (dtype == PFB_F64 || (flags & PFB_ARITH_F64)) ? 8 : 4;
    const long long Q = (long long)C * b->B;
    const size_t nstate = (size_t)Q * b->K * 2;
    if (!b->per_channel && x != y) {
        // whole-signal channel groups when one channel's band signals fit the device block (16 GiB per slot by default, at most a quarter of the free device memory)
        const long long per_channel_bytes = (long long)b->B * ((n + 31) / 32 * 32) * (long long)esz;
        long long cap_bytes = (long long)env_int("PFB_HOST_BLOCK_MB", 16384) << 20;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)   // two slots; leave most of the device to the caller
            cap_bytes = std::min<long long>(cap_bytes, (long long)((free_b + b->hy_bytes * 2) / 4));
        if (per_channel_bytes <= cap_bytes) {
            // at least four groups when there are that many channels, so input copy, kernels and output drain overlap
            long long Cg = std::max<long long>(1, std::min<long long>(C, cap_bytes / per_channel_bytes));
            Cg = std::min<long long>(Cg, std::max<long long>(1, (C + 3) / 4));
            return filter_host_by_channels(b, dtype, x, ldx, y, ldy, states, n, C, flags, (int)Cg);
        }
    }
    // time block: multiple of 32 elements, sized so one output block is at most ~1 GiB
    long long nb = n;
    const long long cap = std::max<long long>(4096, (1LL << 30) / (long long)(esz * Q));
    if (nb > cap) nb = cap / 32 * 32;
    if (nb < 1) nb = 1;
    const long long pitch = (nb + 31) / 32 * 32;

    void *dx = nullptr, *dy[2] = {nullptr, nullptr}, *ds = nullptr;
    if (int rc0 = b->ensure_host_ctx()) return rc0;
    cudaStream_t sc = b->hs[0], sd = b->hs[1];
    cudaEvent_t done[2] = {b->hev[0], b->hev[1]}, drained[2] = {b->hev[2], b->hev[3]};
    int rc = PFB_OK;
    auto cleanup = [&]() {
        cudaStreamSynchronize(sc);
        cudaStreamSynchronize(sd);
        cudaFree(dx);
        cudaFree(dy[0]);
        cudaFree(dy[1]);
        cudaFree(ds);
    };
#define PFB_CUDA_H(call)                                                                                  \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            rc = fail(PFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            cleanup();                                                                                    \
            return rc;                                                                                    \
        }                                                                                                 \
    } while (0)
    // two input blocks [C][pitch] and two output blocks [Q][pitch]
    PFB_CUDA_H(counted_malloc(&dx, 2 * (size_t)C * pitch * esz));
    PFB_CUDA_H(counted_malloc(&dy[0], (size_t)Q * pitch * esz));
    if (n > nb) PFB_CUDA_H(counted_malloc(&dy[1], (size_t)Q * pitch * esz));
    PFB_CUDA_H(counted_malloc(&ds, nstate * ssz));
    PFB_CUDA_H(cudaMemcpyAsync(ds, states, nstate * ssz, cudaMemcpyHostToDevice, sc));

    const long long nblocks = n == 0 ? 0 : (n + nb - 1) / nb;
    for (long long blk = 0; blk < nblocks; ++blk) {
        const int s = (int)(blk & 1);
        const long long t0 = blk * nb, len = std::min(nb, n - t0);
        char *dxs = (char *)dx + (size_t)s * C * pitch * esz;
        // input block: copied on the drain stream's sibling (compute stream orders it before the kernel)
        PFB_CUDA_H(cudaMemcpy2DAsync(dxs, pitch * esz, (const char *)x + t0 * esz, ldx * esz, len * esz, C,
                                     cudaMemcpyHostToDevice, sc));
        if (blk >= 2) PFB_CUDA_H(cudaStreamWaitEvent(sc, drained[s], 0));   // output slot free again
        rc = pfb_bank_filter(b, dtype, dxs, pitch, dy[s], pitch, ds, len, C, flags, sc);
        if (rc) {
            cleanup();
            return rc;
        }
        PFB_CUDA_H(cudaEventRecord(done[s], sc));
        PFB_CUDA_H(cudaStreamWaitEvent(sd, done[s], 0));
        // output rows q = c*B + b go to y[q][t0 .. t0+len)
        PFB_CUDA_H(cudaMemcpy2DAsync((char *)y + t0 * esz, ldy * esz, dy[s], pitch * esz, len * esz, (size_t)Q,
                                     cudaMemcpyDeviceToHost, sd));
        PFB_CUDA_H(cudaEventRecord(drained[s], sd));
    }
    PFB_CUDA_H(cudaMemcpyAsync(states, ds, nstate * ssz, cudaMemcpyDeviceToHost, sc));
    PFB_CUDA_H(cudaStreamSynchronize(sc));
    PFB_CUDA_H(cudaStreamSynchronize(sd));
    cleanup();
    g_status = PFB_OK;
    return PFB_OK;
#undef PFB_CUDA_H
}

// ------------------------------------------------------------------------------------------------
// Weighted bank: SPL-weighting prefilter (splweighting.py:61-80) + bank (octbank.py:412-434) as ONE call.
// ------------------------------------------------------------------------------------------------
int pfb_bank_filter_weighted(pfb_bank *b, int dtype, const void *x, int64_t ldx, void *y, int64_t ldy, void *states,
                             const double *wb, int nwb, const double *wa, int nwa, double *wz, int64_t n, int C, int flags,
                             void *stream_) {
    if (!b) return fail(PFB_ERR_INVALID, "pfb_bank_filter_weighted: null bank");
    if (n == 0) {
        g_status = PFB_OK;
        return PFB_OK;
    }
    if (!x || !y || !states || !wb || !wa || !wz || nwb < 1 || nwa < 1 || nwb > 8 || nwa > 8 || C <= 0 || n < 0 || ldx < n)
        return fail(PFB_ERR_INVALID, "pfb_bank_filter_weighted: bad argument (at most 8 taps)");
    if (dtype != PFB_F32 && dtype != PFB_F64) return fail(PFB_ERR_INVALID, "pfb_bank_filter_weighted: bad dtype %d", dtype);
    if (wa[0] == 0.0) return fail(PFB_ERR_INVALID, "pfb_bank_filter_weighted: a[0] must not be 0");
    if (flags & (PFB_DECIMATE2 | PFB_DEC_PHASE1)) return fail(PFB_ERR_INVALID, "pfb_bank_filter_weighted: no decimation here");
    if ((const void *)x == (const void *)y) return fail(PFB_ERR_INVALID, "pfb_bank_filter_weighted: x must not alias y");
    std::lock_guard<std::mutex> lock(b->mu);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int ntaps = std::max(nwb, nwa), nz = std::max(ntaps - 1, 1);
    const size_t esz = dtype == PFB_F64 ? 8 : 4;
    FilterExtra ex;
    for (int k = 0; k < 8; ++k) {   // scipy normalises by a[0] first
        ex.wb[k] = k < nwb ? wb[k] / wa[0] : 0.0;
        ex.wa[k] = k < nwa ? wa[k] / wa[0] : 0.0;
    }
    ex.wnz = nz;
    ex.wz = wz;
    // Fused pass: one chunk per cascade, so it pays when (channel blocks x band groups) fill a good part of the
    // machine; few channels take the weighting kernel and the time-chunked bank.  PFB_WEIGHT_FUSED=0/1 overrides.
    const int G = (b->B + pfb::kTmaWeightBands - 1) / pfb::kTmaWeightBands;
    bool fused = !(flags & PFB_EXACT) && nz >= 4 && nz <= 6 && (long long)((C + 31) / 32) * G * 2 >= b->sm_count;
    fused = env_int("PFB_WEIGHT_FUSED", fused ? 1 : 0) != 0;
    long long done = 0;
    int kernels = 0;
    pfb_launch_info fused_info{};
    if (fused) {
        int rc = bank_filter_locked(b, dtype, x, ldx, y, ldy, states, n, C, flags, stream_, &ex);
        if (rc < 0) return rc;
        if (rc == PFB_OK) {
            done = ex.n_done;
            kernels += b->last.kernels;
            fused_info = b->last;
        }
    }
    if (done < n) {   // unfused (remainder of the) signal: weighting kernel into a scratch row set, then the bank
        const long long rem = n - done;
        const long long pitch = (rem + 31) / 32 * 32;
        void *tmp = nullptr;
        PFB_CUDA(counted_malloc_async(&tmp, (size_t)C * pitch * esz, stream));
        pfb::TfParams tp{};
        tp.x = (const char *)x + done * esz, tp.y = tmp, tp.z = wz;
        tp.n = rem, tp.ldx = ldx, tp.ldy = pitch, tp.R = C, tp.ntaps = ntaps;
        for (int k = 0; k < 8; ++k) tp.b[k] = ex.wb[k], tp.a[k] = ex.wa[k];
        const bool al = aligned16(tp.x) && (ldx * esz) % 16 == 0;
        int err = pfb::launch_tf(tp, dtype, al, stream);
        int rc = PFB_OK;
        if (err) rc = fail(PFB_ERR_CUDA, "weighting kernel launch failed: %s", cudaGetErrorString((cudaError_t)err));
        ++kernels;
        if (!rc) rc = bank_filter_locked(b, dtype, tmp, pitch, (char *)y + done * esz, ldy, states, rem, C, flags, stream_);
        cudaFreeAsync(tmp, stream);
        if (rc) return rc;
        kernels += b->last.kernels;
    }
    if (done > 0) b->last = fused_info;
    b->last.kernels = kernels;
    g_status = PFB_OK;
    return PFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Decimating band path as one call: stage m runs at fs / 2^m; every stage's bands and the anti-alias
// low-pass that feeds the next stage are filtered in ONE pass over the stage signal (the anti-alias
// cascade is one more band warp whose output is stored decimated), so no band signal is ever written
// at a rate above its own and the Python loop over stages is gone.
// ------------------------------------------------------------------------------------------------
}  // extern "C"

struct pfb_decbank {
    int device = 0;
    int M = 0;                         // halving stages; stages 0 .. M
    int K = 0, Kaa = 0, Kp = 0;        // band sections, anti-alias sections, sections of the padded anti-alias cascade
    bool fusable = false;              // Kp == K: the anti-alias cascade fits the bands' kernel as one more band
    std::vector<int> nb;               // bands per stage
    std::vector<pfb_bank *> fused;     // stage m < M: bands + anti-alias cascade (nb[m] + 1 rows), or null
    std::vector<pfb_bank *> bands;     // stage m: nb[m] rows, or null when the stage has no band
    pfb_bank *aa = nullptr;            // the anti-alias cascade alone (1 row, Kp sections)
    void *sig[2] = {nullptr, nullptr}; // stage signals (ping-pong): stage m reads sig[m & 1] (m > 0), writes sig[(m + 1) & 1]
    size_t sig_bytes[2] = {0, 0};
    // split schedule (split_filter): every stage signal has its own buffer, the anti-alias chain its own stream, every
    // stage's bands theirs
    std::vector<void *> ssig;          // ssig[m]: signal of stage m (m >= 1)
    std::vector<size_t> ssig_bytes;
    cudaStream_t s_aa = nullptr;
    std::vector<cudaStream_t> s_band;  // s_band[m], m >= 1 (stage 0 runs on the caller's stream)
    cudaEvent_t ev_start = nullptr;
    std::vector<cudaEvent_t> ev_aa, ev_band;
    int ensure_split(int C, long long n);
    int split_filter(const double *x, int64_t ldx, void *const *y, const int64_t *ldy, void *const *band_states,
                     double *aa_states, int *phases, int64_t n, int C, int flags, cudaStream_t stream);
    std::mutex mu;
    int last_kernels = 0, last_fused = 0, last_split = 0;
    long long split_calls = 0;
};

// ------------------------------------------------------------------------------------------------
// Split schedule of the decimating band path.  The stage-by-stage schedule (below, pfb_decbank_filter) runs nine
// dependent launch sets; stages 1 .. M hold few cascades each (13 of 119 bands per octave), so each of them fills the
// machine badly and the next one cannot start before its last CTA has finished.  Only the anti-alias CHAIN is
// sequential, though: stage m's bands need nothing but the stage signal m.  Here the chain (one cascade per channel and
// stage, stored decimated by the same DEC kernels) runs on a high-priority stream of its own, stage 0's bands on the
// caller's stream from the start, and the bands of stage m >= 1 on stream m as soon as signal m exists -- the small
// stages overlap each other and the tail of stage 0.  Same kernels, same arithmetic, same stores: every band signal is
// still written once at its own rate, and the anti-alias outputs only decimated.
// ------------------------------------------------------------------------------------------------
int pfb_decbank::ensure_split(int C, long long n) {
    if (!s_aa) {
        int lo = 0, hi = 0;
        PFB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PFB_CUDA(cudaStreamCreateWithPriority(&s_aa, cudaStreamNonBlocking, hi));
        PFB_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
        s_band.assign(M + 1, nullptr);
        ev_aa.assign(M + 1, nullptr);
        ev_band.assign(M + 1, nullptr);
        for (int m = 0; m <= M; ++m) {
            if (m >= 1) PFB_CUDA(cudaStreamCreateWithPriority(&s_band[m], cudaStreamNonBlocking, hi));
            PFB_CUDA(cudaEventCreateWithFlags(&ev_aa[m], cudaEventDisableTiming));
            PFB_CUDA(cudaEventCreateWithFlags(&ev_band[m], cudaEventDisableTiming));
        }
        ssig.assign(M + 1, nullptr);
        ssig_bytes.assign(M + 1, 0);
    }
    long long nm = n;
    for (int m = 1; m <= M; ++m) {
        nm = (nm + 1) / 2;
        const long long pitch = (nm + 31) / 32 * 32;
        const size_t need = (size_t)C * (size_t)pitch * sizeof(double);
        if (ssig_bytes[m] < need) {
            if (ssig[m]) {
                PFB_CUDA(cudaDeviceSynchronize());
                cudaFree(ssig[m]);
                ssig[m] = nullptr;
                ssig_bytes[m] = 0;
            }
            PFB_CUDA(counted_malloc(&ssig[m], need));
            ssig_bytes[m] = need;
        }
    }
    return PFB_OK;
}

int pfb_decbank::split_filter(const double *x, int64_t ldx, void *const *y, const int64_t *ldy, void *const *band_states,
                              double *aa_states, int *phases, int64_t n, int C, int flags, cudaStream_t stream) {
    if (int rc = ensure_split(C, n)) return rc;
    last_kernels = last_fused = 0;
    last_split = 1;
    // PFB_DEC_SPLIT_TRACE=1: when each chain link and each stage finished, relative to the start of the call (stderr)
    const bool trace = env_int("PFB_DEC_SPLIT_TRACE", 0) != 0;
    std::vector<cudaEvent_t> tr_aa(M + 1, nullptr), tr_band(M + 1, nullptr);
    cudaEvent_t tr0 = nullptr;
    auto tr_rec = [&](cudaEvent_t &e, cudaStream_t st) {
        if (!trace) return;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
    };
    tr_rec(tr0, stream);
    PFB_CUDA(cudaEventRecord(ev_start, stream));
    PFB_CUDA(cudaStreamWaitEvent(s_aa, ev_start, 0));
    // a previous call (possibly issued from another stream) may still be reading the stage signals this one overwrites
    if (split_calls++ > 0)
        for (int m = 1; m <= M; ++m) PFB_CUDA(cudaStreamWaitEvent(s_aa, ev_band[m], 0));
    // 1. the anti-alias chain: stage signal m -> m + 1, decimated stores only
    std::vector<long long> len(M + 1), pitch(M + 1);
    len[0] = n, pitch[0] = ldx;
    {
        const double *in = x;
        long long nm = n;
        for (int m = 0; m < M && nm > 0; ++m) {
            const int ph = phases[m] & 1;
            const long long n_next = (nm - ph + 1) / 2;
            const long long pitch2 = ((nm + 1) / 2 + 31) / 32 * 32;
            double *out2 = (double *)ssig[m + 1];
            double *aast = aa_states + (size_t)m * C * Kp * 2;
            long long done = 0;
            {
                FilterExtra ex;
                ex.dec = true, ex.dec_phase = ph, ex.y2 = out2, ex.ldy2 = pitch2, ex.dec_states = aast;
                std::lock_guard<std::mutex> block(aa->mu);
                int rc = bank_filter_locked(aa, PFB_F64, in, pitch[m], nullptr, (nm + 31) / 32 * 32, (void *)aast, nm, C, flags,
                                            (void *)s_aa, &ex);
                if (rc < 0) return rc;
                if (rc == PFB_OK && ex.n_done > 0) {
                    done = ex.n_done;
                    last_kernels += aa->last.kernels;
                    ++last_fused;
                }
            }
            if (done < nm) {
                int rc = pfb_bank_filter(aa, PFB_F64, in + done, pitch[m], out2 + done / 2, pitch2, aast, nm - done, C,
                                         flags | PFB_DECIMATE2 | (ph ? PFB_DEC_PHASE1 : 0), (void *)s_aa);
                if (rc) return rc;
                last_kernels += aa->last.kernels;
            }
            PFB_CUDA(cudaEventRecord(ev_aa[m], s_aa));
            tr_rec(tr_aa[m], s_aa);
            phases[m] = (int)((ph + nm) & 1);
            in = out2;
            len[m + 1] = n_next, pitch[m + 1] = pitch2;
            nm = n_next;
        }
    }
    // 2. the bands of every stage, each stage on its own stream behind the chain link that produces its signal
    for (int m = 0; m <= M; ++m) {
        if (nb[m] <= 0 || len[m] <= 0) {
            if (m > 0) PFB_CUDA(cudaEventRecord(ev_band[m], s_band[m]));
            continue;
        }
        if (m > 0 && len[m] > 0 && (!y[m] || !band_states[m] || ldy[m] < len[m]))
            return fail(PFB_ERR_INVALID, "pfb_decbank_filter: stage %d needs an output of %lld samples per row", m, len[m]);
        cudaStream_t st = m == 0 ? stream : s_band[m];
        if (m > 0) PFB_CUDA(cudaStreamWaitEvent(st, ev_aa[m - 1], 0));
        // The stages 1 .. M run side by side: their chunk counts are chosen as if each had an eighth of the machine
        // (fewer chunks, less pre-pass work; measured 17.4-18.1 ms per 4 s block of cfg 3 against 18.3-18.8 with the
        // whole machine assumed, profiles/r04f_dec_split_sweep.txt, r04i_dec_split_sweep.txt).  PFB_DEC_S0_J: chunks
        // of stage 0 (experiments; 2 .. 16 are within the noise).
        if (m > 0) bands[m]->slot_share = std::max(0.01, std::min(1.0, atof(getenv("PFB_DEC_SPLIT_SHARE") ? getenv("PFB_DEC_SPLIT_SHARE") : "0.125")));
        if (m == 0) bands[m]->force_J = env_int("PFB_DEC_S0_J", 0);
        const double *in = m == 0 ? x : (const double *)ssig[m];
        int rc = pfb_bank_filter(bands[m], PFB_F64, in, pitch[m], y[m], ldy[m], band_states[m], len[m], C, flags, (void *)st);
        if (rc) return rc;
        last_kernels += bands[m]->last.kernels;
        if (m > 0) PFB_CUDA(cudaEventRecord(ev_band[m], st));
        tr_rec(tr_band[m], st);
    }
    // 3. join: the caller's stream continues when the chain and every stage have finished
    if (M > 0) {
        PFB_CUDA(cudaEventRecord(ev_aa[M], s_aa));
        PFB_CUDA(cudaStreamWaitEvent(stream, ev_aa[M], 0));
    }
    for (int m = 1; m <= M; ++m)
        if (nb[m] > 0 && len[m] > 0) PFB_CUDA(cudaStreamWaitEvent(stream, ev_band[m], 0));
    if (trace) {
        cudaStreamSynchronize(stream);
        fprintf(stderr, "[dec split] finished at (ms after the start):");
        for (int m = 0; m <= M; ++m) {
            float a = -1.f, bnd = -1.f;
            if (tr_aa[m]) cudaEventElapsedTime(&a, tr0, tr_aa[m]), cudaEventDestroy(tr_aa[m]);
            if (tr_band[m]) cudaEventElapsedTime(&bnd, tr0, tr_band[m]), cudaEventDestroy(tr_band[m]);
            fprintf(stderr, "  stage %d: chain %.2f bands %.2f (J=%d)", m, a, bnd,
                    bands[m] ? bands[m]->last.warps_per_cascade * std::min<int>(32, std::max(1, (int)(len[m] / std::max<long long>(1, bands[m]->last.chunk)) )) : 0);
        }
        fprintf(stderr, "\n");
        cudaEventDestroy(tr0);
    }
    g_status = PFB_OK;
    return PFB_OK;
}

extern "C" {

void pfb_decbank_destroy(pfb_decbank *d) {
    if (!d) return;
    for (pfb_bank *b : d->fused) pfb_bank_destroy(b);
    for (pfb_bank *b : d->bands) pfb_bank_destroy(b);
    if (d->s_aa) {
        cudaStreamSynchronize(d->s_aa);
        for (cudaStream_t st : d->s_band)
            if (st) cudaStreamSynchronize(st);
    }
    pfb_bank_destroy(d->aa);
    cudaFree(d->sig[0]);
    cudaFree(d->sig[1]);
    for (void *q : d->ssig) cudaFree(q);
    if (d->s_aa) cudaStreamDestroy(d->s_aa);
    for (cudaStream_t st : d->s_band)
        if (st) cudaStreamDestroy(st);
    if (d->ev_start) cudaEventDestroy(d->ev_start);
    for (cudaEvent_t e : d->ev_aa)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : d->ev_band)
        if (e) cudaEventDestroy(e);
    delete d;
}

int pfb_decbank_create(pfb_decbank **out, int nstages, const int *bands_per_stage, const double *const *sos_per_stage, int K,
                       const double *aa_sos, int Kaa) {
    if (!out || nstages < 1 || !bands_per_stage || !sos_per_stage || K <= 0 || (nstages > 1 && (!aa_sos || Kaa <= 0)))
        return fail(PFB_ERR_INVALID, "pfb_decbank_create: bad arguments");
    pfb_decbank *d = new pfb_decbank;
    d->M = nstages - 1;
    d->K = K;
    d->Kaa = nstages > 1 ? Kaa : 0;
    // the decimating store needs a cascade that runs in one sweep: pad to the next sweep size
    int Kp = std::max(K, d->Kaa);
    static const int sizes[] = {1, 2, 3, 4, 6, 8};
    for (int sz : sizes)
        if (sz >= Kp) { Kp = sz; break; }
    if (d->M > 0 && Kp > 8) {
        delete d;
        return fail(PFB_ERR_INVALID, "pfb_decbank_create: cascades of more than 8 sections cannot decimate");
    }
    d->Kp = Kp;
    d->fusable = d->M > 0 && Kp == K;
    d->nb.assign(bands_per_stage, bands_per_stage + nstages);
    // anti-alias cascade padded with identity sections (b0 = 1: they pass their input through unchanged)
    std::vector<double> aap((size_t)Kp * 6, 0.0);
    for (int k = 0; k < Kp; ++k) {
        if (k < d->Kaa) std::copy(aa_sos + 6 * k, aa_sos + 6 * k + 6, &aap[6 * k]);
        else aap[6 * k] = 1.0, aap[6 * k + 3] = 1.0;
    }
    int rc = PFB_OK;
    PFB_CUDA(cudaGetDevice(&d->device));
    if (d->M > 0) rc = pfb_bank_create(&d->aa, aap.data(), 1, 1, Kp, 0);
    d->fused.assign(nstages, nullptr);
    d->bands.assign(nstages, nullptr);
    for (int m = 0; m < nstages && rc == PFB_OK; ++m) {
        const int B = d->nb[m];
        if (B < 0 || (B > 0 && !sos_per_stage[m])) rc = fail(PFB_ERR_INVALID, "pfb_decbank_create: stage %d has no coefficients", m);
        if (rc == PFB_OK && B > 0) rc = pfb_bank_create(&d->bands[m], sos_per_stage[m], 1, B, K, 0);
        if (rc == PFB_OK && m < d->M && d->fusable) {
            std::vector<double> f((size_t)(B + 1) * K * 6);
            if (B > 0) std::copy(sos_per_stage[m], sos_per_stage[m] + (size_t)B * K * 6, f.begin());
            std::copy(aap.begin(), aap.end(), f.begin() + (size_t)B * K * 6);
            rc = pfb_bank_create(&d->fused[m], f.data(), 1, B + 1, K, 0);
        }
    }
    if (rc != PFB_OK) {
        pfb_decbank_destroy(d);
        return rc;
    }
    *out = d;
    g_status = PFB_OK;
    return PFB_OK;
}

int pfb_decbank_info(const pfb_decbank *d, int *stages, int *aa_sections_padded, int *last_kernels, int *last_fused_stages) {
    if (!d) return fail(PFB_ERR_INVALID, "pfb_decbank_info: null argument");
    if (stages) *stages = d->M + 1;
    if (aa_sections_padded) *aa_sections_padded = d->Kp;
    if (last_kernels) *last_kernels = d->last_kernels;
    if (last_fused_stages) *last_fused_stages = d->last_fused;
    return PFB_OK;
}

int pfb_decbank_filter(pfb_decbank *d, const double *x, int64_t ldx, void *const *y, const int64_t *ldy,
                       void *const *band_states, double *aa_states, int *phases, int64_t n, int C, int flags, void *stream_) {
    if (!d || !x || !y || !ldy || !band_states || (d->M > 0 && (!aa_states || !phases)) || n < 0 || C <= 0 || ldx < n)
        return fail(PFB_ERR_INVALID, "pfb_decbank_filter: bad argument");
    if (flags & (PFB_DECIMATE2 | PFB_DEC_PHASE1 | PFB_ARITH_F64))
        return fail(PFB_ERR_INVALID, "pfb_decbank_filter: flags may only select the mode and PFB_EXACT");
    std::lock_guard<std::mutex> lock(d->mu);
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = -1;
    PFB_CUDA(cudaGetDevice(&dev));
    if (dev != d->device) return fail(PFB_ERR_INVALID, "pfb_decbank_filter: created on device %d, current device is %d", d->device, dev);
    // stage signal buffers: stage m + 1 holds at most ceil(n_m / 2) samples per channel
    {
        long long nm = n;
        size_t need[2] = {0, 0};
        for (int m = 0; m < d->M; ++m) {
            nm = (nm + 1) / 2;
            const long long pitch = (nm + 31) / 32 * 32;
            need[(m + 1) & 1] = std::max<size_t>(need[(m + 1) & 1], (size_t)C * (size_t)pitch * sizeof(double));
        }
        for (int i = 0; i < 2; ++i)
            if (d->sig_bytes[i] < need[i]) {
                if (d->sig[i]) {
                    PFB_CUDA(cudaDeviceSynchronize());
                    cudaFree(d->sig[i]);
                    d->sig[i] = nullptr;
                    d->sig_bytes[i] = 0;
                }
                PFB_CUDA(counted_malloc(&d->sig[i], need[i]));
                d->sig_bytes[i] = need[i];
            }
    }
    const bool exact = (flags & PFB_EXACT) != 0;
    const bool try_fused = d->fusable && !exact && !env_int("PFB_DEC_UNFUSED", 0);
    d->last_split = 0;
    if (try_fused && d->M > 0 && n > 0 && env_int("PFB_DEC_SPLIT", 1)) {
        for (int m = 0; m <= d->M; ++m)
            if (d->nb[m] > 0 && m == 0 && (!y[m] || !band_states[m] || ldy[m] < n))
                return fail(PFB_ERR_INVALID, "pfb_decbank_filter: stage %d needs an output of %lld samples per row", m, (long long)n);
        const int rcs = d->split_filter(x, ldx, y, ldy, band_states, aa_states, phases, n, C, flags, stream);
        if (rcs != PFB_OK && d->s_aa) {
            // a failed call must not leave work of its side streams running on buffers the caller may now release
            cudaStreamSynchronize(d->s_aa);
            for (cudaStream_t st : d->s_band)
                if (st) cudaStreamSynchronize(st);
        }
        return rcs;
    }
    d->last_kernels = d->last_fused = 0;
    const double *in = x;
    long long ld_in = ldx, nm = n;
    for (int m = 0; m <= d->M && nm > 0; ++m) {
        const int B = d->nb[m];
        if (B > 0 && (!y[m] || !band_states[m] || ldy[m] < nm))
            return fail(PFB_ERR_INVALID, "pfb_decbank_filter: stage %d needs an output of %lld samples per row", m, nm);
        if (m == d->M) {
            if (B > 0) {
                int rc = pfb_bank_filter(d->bands[m], PFB_F64, in, ld_in, y[m], ldy[m], band_states[m], nm, C, flags, stream_);
                if (rc) return rc;
                d->last_kernels += d->bands[m]->last.kernels;
            }
            break;
        }
        const int ph = phases[m] & 1;
        const long long n_next = (nm - ph + 1) / 2;
        const long long pitch2 = ((nm + 1) / 2 + 31) / 32 * 32;
        double *out2 = (double *)d->sig[(m + 1) & 1];
        double *aast = aa_states + (size_t)m * C * d->Kp * 2;
        long long done = 0;
        if (try_fused) {
            FilterExtra ex;
            ex.dec = true, ex.dec_phase = ph, ex.y2 = out2, ex.ldy2 = pitch2, ex.dec_states = aast;
            pfb_bank *fb = d->fused[m];
            std::lock_guard<std::mutex> block(fb->mu);
            int rc = bank_filter_locked(fb, PFB_F64, in, ld_in, B > 0 ? y[m] : nullptr, B > 0 ? ldy[m] : (nm + 31) / 32 * 32,
                                        B > 0 ? band_states[m] : (void *)aast, nm, C, flags, stream_, &ex);
            if (rc < 0) return rc;
            if (rc == PFB_OK && ex.n_done > 0) {
                done = ex.n_done;
                d->last_kernels += fb->last.kernels;
                ++d->last_fused;
            }
        }
        if (done < nm) {   // (ragged end of) the stage: bands and the anti-alias cascade as separate passes
            const long long rem = nm - done;
            if (B > 0) {
                int rc = pfb_bank_filter(d->bands[m], PFB_F64, in + done, ld_in, (double *)y[m] + done, ldy[m], band_states[m], rem, C,
                                         flags, stream_);
                if (rc) return rc;
                d->last_kernels += d->bands[m]->last.kernels;
            }
            // `done` is even, so the remainder starts on the decimator phase of the block and its kept samples
            // continue at output index done / 2
            int rc = pfb_bank_filter(d->aa, PFB_F64, in + done, ld_in, out2 + done / 2, pitch2, aast, rem, C,
                                     flags | PFB_DECIMATE2 | (ph ? PFB_DEC_PHASE1 : 0), stream_);
            if (rc) return rc;
            d->last_kernels += d->aa->last.kernels;
        }
        phases[m] = (int)((ph + nm) & 1);
        in = out2;
        ld_in = pitch2;
        nm = n_next;
    }
    g_status = PFB_OK;
    return PFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Streaming session: block-wise filtering of HOST audio with device-resident states (the reason the
// reference's filter functions carry states: blockprocessing.py:3-72, sosfiltering.py:143-147).
// Streams, events and staging slots are created once; pushing a block enqueues H2D, kernels and D2H
// and returns; nothing is allocated after the first block of a given length.
// ------------------------------------------------------------------------------------------------
}  // extern "C"

struct pfb_stream {
    pfb_bank *bank = nullptr;
    int device = 0, dtype = PFB_F64, C = 0, flags = 0;
    long long max_block = 0, pitch = 0;
    size_t esz = 8, ssz = 8;
    cudaStream_t sh = nullptr, sc = nullptr, sd = nullptr;      // input copies, kernels, output copies
    cudaEvent_t in_done[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
    void *dx[2] = {nullptr, nullptr}, *dy[2] = {nullptr, nullptr}, *ds = nullptr;
    long long blocks = 0, samples = 0;
    long long allocs_at_create = 0, allocs_after_first = -1;
    std::mutex mu;
};

extern "C" {

void pfb_stream_destroy(pfb_stream *s) {
    if (!s) return;
    if (s->sc) {
        cudaStreamSynchronize(s->sh);
        cudaStreamSynchronize(s->sc);
        cudaStreamSynchronize(s->sd);
        pfb_bank_release_stream(s->bank, (void *)s->sc);      // the bank's per-stream scratch goes with the stream
    }
    for (int i = 0; i < 2; ++i) {
        cudaFree(s->dx[i]);
        cudaFree(s->dy[i]);
        if (s->in_done[i]) cudaEventDestroy(s->in_done[i]);
        if (s->done[i]) cudaEventDestroy(s->done[i]);
        if (s->drained[i]) cudaEventDestroy(s->drained[i]);
    }
    cudaFree(s->ds);
    if (s->sh) cudaStreamDestroy(s->sh);
    if (s->sc) cudaStreamDestroy(s->sc);
    if (s->sd) cudaStreamDestroy(s->sd);
    delete s;
}

int pfb_stream_create(pfb_stream **out, pfb_bank *bank, int dtype, int C, int64_t max_block, int flags) {
    if (!out || !bank || C <= 0 || max_block <= 0 || (dtype != PFB_F32 && dtype != PFB_F64))
        return fail(PFB_ERR_INVALID, "pfb_stream_create: bad arguments");
    if (flags & (PFB_DECIMATE2 | PFB_DEC_PHASE1)) return fail(PFB_ERR_INVALID, "pfb_stream_create: no decimation here");
    if (bank->per_channel && C != bank->C_sos) return fail(PFB_ERR_INVALID, "pfb_stream_create: bank holds %d channels", bank->C_sos);
    if (int rc = check_device(bank, "pfb_stream_create")) return rc;
    pfb_stream *s = new pfb_stream;
    s->bank = bank;
    s->device = bank->device;
    if (flags & PFB_ARITH_AUTO) flags |= PFB_ARITH_F64;   // per-band arithmetic implies float64 states
    s->dtype = dtype, s->C = C, s->flags = flags;
    s->max_block = max_block;
    s->pitch = (max_block + 31) / 32 * 32;
    s->esz = dtype == PFB_F64 ? 8 : 4;
    s->ssz = (dtype == PFB_F64 || (flags & PFB_ARITH_F64)) ? 8 : 4;
    s->allocs_at_create = g_device_allocs.load();
    const size_t Q = (size_t)C * bank->B;
    cudaError_t e = cudaStreamCreateWithFlags(&s->sh, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->sc, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->sd, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&s->in_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->drained[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = counted_malloc(&s->dx[i], (size_t)C * s->pitch * s->esz);
        if (e == cudaSuccess) e = counted_malloc(&s->dy[i], Q * s->pitch * s->esz);
    }
    const size_t state_bytes = Q * bank->K * 2 * s->ssz;
    if (e == cudaSuccess) e = counted_malloc(&s->ds, state_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->ds, 0, state_bytes, s->sc);
    if (e != cudaSuccess) {
        pfb_stream_destroy(s);
        return fail(PFB_ERR_CUDA, "pfb_stream_create: %s", cudaGetErrorString(e));
    }
    *out = s;
    g_status = PFB_OK;
    return PFB_OK;
}

int pfb_stream_push(pfb_stream *s, const void *x, int64_t ldx, void *y, int64_t ldy, int64_t n) {
    if (!s || !x || !y || n < 0 || n > s->max_block || ldx < n || ldy < n)
        return fail(PFB_ERR_INVALID, "pfb_stream_push: bad argument (block of at most %lld samples)", s ? s->max_block : 0LL);
    if (n == 0) {
        g_status = PFB_OK;
        return PFB_OK;
    }
    std::lock_guard<std::mutex> lock(s->mu);
    if (int rc = check_device(s->bank, "pfb_stream_push")) return rc;
    const int slot = (int)(s->blocks & 1);
    const size_t esz = s->esz, Q = (size_t)s->C * s->bank->B;
    // the slot's previous block (two pushes ago) must have left the device before its buffers are reused
    if (s->blocks >= 2) PFB_CUDA(cudaStreamWaitEvent(s->sh, s->drained[slot], 0));
    PFB_CUDA(cudaMemcpy2DAsync(s->dx[slot], s->pitch * esz, x, ldx * esz, n * esz, (size_t)s->C, cudaMemcpyHostToDevice, s->sh));
    PFB_CUDA(cudaEventRecord(s->in_done[slot], s->sh));
    PFB_CUDA(cudaStreamWaitEvent(s->sc, s->in_done[slot], 0));
    if (s->blocks >= 2) PFB_CUDA(cudaStreamWaitEvent(s->sc, s->drained[slot], 0));
    int rc = pfb_bank_filter(s->bank, s->dtype, s->dx[slot], s->pitch, s->dy[slot], s->pitch, s->ds, n, s->C, s->flags, (void *)s->sc);
    if (rc) return rc;
    PFB_CUDA(cudaEventRecord(s->done[slot], s->sc));
    PFB_CUDA(cudaStreamWaitEvent(s->sd, s->done[slot], 0));
    PFB_CUDA(cudaMemcpy2DAsync(y, ldy * esz, s->dy[slot], s->pitch * esz, n * esz, Q, cudaMemcpyDeviceToHost, s->sd));
    PFB_CUDA(cudaEventRecord(s->drained[slot], s->sd));
    ++s->blocks;
    s->samples += n;
    if (s->allocs_after_first < 0) s->allocs_after_first = g_device_allocs.load();
    g_status = PFB_OK;
    return PFB_OK;
}

int pfb_stream_sync(pfb_stream *s) {
    if (!s) return fail(PFB_ERR_INVALID, "pfb_stream_sync: null session");
    PFB_CUDA(cudaStreamSynchronize(s->sh));
    PFB_CUDA(cudaStreamSynchronize(s->sc));
    PFB_CUDA(cudaStreamSynchronize(s->sd));
    g_status = PFB_OK;
    return PFB_OK;
}

int pfb_stream_wait(pfb_stream *s, int64_t block) {
    if (!s || block < 0) return fail(PFB_ERR_INVALID, "pfb_stream_wait: bad argument");
    cudaEvent_t ev = nullptr;
    {
        std::lock_guard<std::mutex> lock(s->mu);
        if (block >= s->blocks) return fail(PFB_ERR_INVALID, "pfb_stream_wait: block %lld has not been pushed", (long long)block);
        // the slot's event belongs to the latest block of the same parity: waiting for it also covers `block`
        ev = s->drained[block & 1];
    }
    PFB_CUDA(cudaEventSynchronize(ev));
    g_status = PFB_OK;
    return PFB_OK;
}

int pfb_stream_states(pfb_stream *s, void *states_host, int set) {
    if (!s || !states_host) return fail(PFB_ERR_INVALID, "pfb_stream_states: null argument");
    std::lock_guard<std::mutex> lock(s->mu);
    const size_t bytes = (size_t)s->C * s->bank->B * s->bank->K * 2 * s->ssz;
    if (set) PFB_CUDA(cudaMemcpyAsync(s->ds, states_host, bytes, cudaMemcpyHostToDevice, s->sc));
    else PFB_CUDA(cudaMemcpyAsync(states_host, s->ds, bytes, cudaMemcpyDeviceToHost, s->sc));
    PFB_CUDA(cudaStreamSynchronize(s->sc));
    g_status = PFB_OK;
    return PFB_OK;
}

int pfb_stream_stats(pfb_stream *s, pfb_stream_info *info) {
    if (!s || !info) return fail(PFB_ERR_INVALID, "pfb_stream_stats: null argument");
    std::lock_guard<std::mutex> lock(s->mu);
    info->blocks = s->blocks;
    info->samples = s->samples;
    const long long now = g_device_allocs.load();
    info->device_allocs_total = now - s->allocs_at_create;
    info->device_allocs_after_first_block = s->allocs_after_first < 0 ? 0 : now - s->allocs_after_first;
    return PFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Reference drop-ins (host pointers, in place).  PFB_DROPIN=fast selects the time-parallel path
// (within 1e-9 relative L2 of the reference in float64); the default is the bit-identical one.
// ------------------------------------------------------------------------------------------------
static void dropin(int dtype, void *signal, long long n, int rows, const void *sos, int K, void *states,
                   const char *who) {
    std::vector<double> s((size_t)rows * K * 6);
    if (dtype == PFB_F64) {
        memcpy(s.data(), sos, s.size() * sizeof(double));
    } else {
        const float *f = (const float *)sos;
        for (size_t i = 0; i < s.size(); ++i) s[i] = f[i];
    }
    pfb_bank *bank = nullptr;
    int rc = pfb_bank_create(&bank, s.data(), rows, 1, K, 1);
    if (rc == PFB_OK) {
        const char *m = getenv("PFB_DROPIN");
        const int flags = (m && !strcmp(m, "fast")) ? PFB_MODE_AUTO : (PFB_MODE_SEQ | PFB_EXACT);
        rc = pfb_bank_filter_host(bank, dtype, signal, n, signal, n, states, n, rows, flags);
    }
    pfb_bank_destroy(bank);
    if (rc != PFB_OK) fprintf(stderr, "libpfb_sos: %s failed (%d): %s\n", who, rc, pfb_last_error());
}

void sosfilter(float *signal, int nsamp, float *sos, int ksos, float *states) {
    dropin(PFB_F32, signal, nsamp, 1, sos, ksos, states, "sosfilter");
}

void sosfilter_double(double *signal, int nsamp, double *sos, int ksos, double *states) {
    dropin(PFB_F64, signal, nsamp, 1, sos, ksos, states, "sosfilter_double");
}

void sosfilter_double_mimo(double *signal, int nframes, int nchan, double *sos, int ksos, int kbands,
                           double *states) {
    // every [c][b] row of `signal` is filtered in place with its own coefficient row
    dropin(PFB_F64, signal, nframes, nchan * kbands, sos, ksos, states, "sosfilter_double_mimo");
}

}  // extern "C"
