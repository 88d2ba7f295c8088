// pfb_inst.inl -- kernel instantiations and launchers for one (signal type, arithmetic type) pair.
// Included by pfb_inst_{dd,ff,fd}.cu with PFB_T, PFB_A and PFB_SUFFIX defined.
#include <cstdlib>
#include <vector>

#include "pfb_kernels.cuh"
#include "pfb_tma.cuh"

#define PFB_CAT2(a, b) a##b
#define PFB_CAT(a, b) PFB_CAT2(a, b)

namespace pfb {

template <int K, bool EXACT, bool ALIGNED>
static int run_seq_a(const SosParams &p, cudaStream_t s) {
    auto kern = sos_seq_kernel<PFB_T, PFB_A, K, EXACT, ALIGNED>;
    const int smem = kSeqWarps * 2 * kTileBytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const int threads = kSeqWarps * 32;
    const unsigned grid = (unsigned)((p.Q + threads - 1) / threads);
    kern<<<grid, threads, smem, s>>>(p);
    return (int)cudaGetLastError();
}
template <int K, bool EXACT>
static int run_seq(const SosParams &p, cudaStream_t s) {
    return p.aligned ? run_seq_a<K, EXACT, true>(p, s) : run_seq_a<K, EXACT, false>(p, s);
}

template <int K, bool ALIGNED>
static int run_scan_a(SosParams p, int warps, const double *coef_host, int rows, int C, cudaStream_t s, int *launches) {
    auto kern = sos_scan_kernel<PFB_T, PFB_A, K, ALIGNED>;
    const int smem = warps * 2 * kTileBytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    static thread_local CoefTable<PFB_A> tab;   // 30 KB: keep it off the stack
    const int rows_per_launch = CoefTable<PFB_A>::kMaxSections / K;
    for (int row0 = 0; row0 < rows; row0 += rows_per_launch) {
        const int nrows = rows - row0 < rows_per_launch ? rows - row0 : rows_per_launch;
        for (int r = 0; r < nrows; ++r)
            for (int k = 0; k < K; ++k)
                for (int i = 0; i < 5; ++i)
                    tab.c[r * K + k][i] = (PFB_A)coef_host[((size_t)(row0 + r) * p.Ktot + p.k0 + k) * 5 + i];
        p.row0 = row0;
        p.nrows = nrows;
        const unsigned grid = p.per_channel ? (unsigned)nrows : (unsigned)nrows * (unsigned)C;
        kern<<<grid, warps * 32, smem, s>>>(p, tab);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        ++*launches;
    }
    return 0;
}
template <int K>
static int run_scan(const SosParams &p, int warps, const double *coef_host, int rows, int C, cudaStream_t s, int *launches) {
    return p.aligned ? run_scan_a<K, true>(p, warps, coef_host, rows, C, s, launches)
                     : run_scan_a<K, false>(p, warps, coef_host, rows, C, s, launches);
}

// Coefficient table of the TMA kernels (kernel parameter block -> uniform registers): [band][section][5]
template <int K, bool SCALED>
static void fill_tma_table(const TmaParams &p, const double *coef_host, CoefTable<PFB_A> &tab) {
    for (int r = 0; r < p.B; ++r) {
        double total_gain = 1.0;
        for (int k = 0; k < K; ++k) {
            const double *c = &coef_host[((size_t)r * p.Ktot + p.k0 + k) * 5];
            for (int i = 0; i < 5; ++i) tab.c[r * K + k][i] = (PFB_A)c[i];
            if (SCALED) {
                // numerator normalised by b0 and written on the OLD states: with w0 = x - a1 w1 - a2 w2,
                // y = w0 + (b1/b0) w1 + (b2/b0) w2 = x + (b1/b0 - a1) w1 + (b2/b0 - a2) w2 (scaled_step).  For the
                // band passes' (1 +- z^-1)^2 numerators the two differences are exact in double (Sterbenz).
                tab.c[r * K + k][1] = (PFB_A)((long double)c[1] / (long double)c[0] - (long double)c[3]);
                tab.c[r * K + k][2] = (PFB_A)((long double)c[2] / (long double)c[0] - (long double)c[4]);
                total_gain *= c[0];
            }
        }
        if (SCALED) tab.c[r * K][0] = (PFB_A)total_gain;
    }
}

// Fused weighting prefilter + bank (sos_tma_weighted_kernel): one chunk per cascade, <= 9 bands per CTA, one launch.
template <int K, bool SCALED>
static int run_tma_weighted(const TmaParams &p, const double *coef_host, const CUtensorMap &xmap, const CUtensorMap &ymap,
                            cudaStream_t s, int *launches, cudaEvent_t *ev, int phases) {
    if (p.J != 1 || p.nb > kTmaWeightBands || p.wide || p.dec_band1) return (int)cudaErrorInvalidValue;
    static thread_local CoefTable<PFB_A> tab;
    fill_tma_table<K, SCALED>(p, coef_host, tab);
    const unsigned grid = (unsigned)p.CB * (unsigned)p.G * (unsigned)p.NW;
    const int last_band_warp = (p.nb - 1) + (p.nb - 1) / 3;              // band slot -> warp index (every fourth is skipped)
    const int threads = 32 * (last_band_warp + 1 > 8 ? last_band_warp + 1 : 8);
    const int smem_in = kTmaWeightStages * 2 * kTileBytes + 3 * kTmaWeightStages * 8;
    if (ev && (phases & 1)) {   // a single chunk per cascade needs no pre-pass
        cudaEventRecord(ev[0], s);
        cudaEventRecord(ev[1], s);
    }
    if (!(phases & 2)) return 0;
    if (ev) cudaEventRecord(ev[2], s);
    cudaError_t e;
    if (p.levels) {
        auto kern = sos_tma_weighted_kernel<PFB_T, PFB_A, K, 2, SCALED>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_in);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, threads, smem_in, s>>>(p, tab, xmap, ymap);
    } else {
        const int smem = smem_in + 2 * p.nb * out_tile_bytes<PFB_T, PFB_A>(2, false, true);
        auto kern = sos_tma_weighted_kernel<PFB_T, PFB_A, K, 1, SCALED>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, threads, smem, s>>>(p, tab, xmap, ymap);
    }
    if (ev) cudaEventRecord(ev[3], s);
    *launches += 1;
    return (int)cudaGetLastError();
}

template <int K, int NBMAX, bool SCALED>
static int run_tma_ns(const TmaParams &p, const double *coef_host, const CUtensorMap &xmap, const CUtensorMap &ymap,
                      const CUtensorMap &ymap2, cudaStream_t s, int *launches, cudaEvent_t *ev, int phases) {
    auto kern_a = sos_tma_kernel<PFB_T, PFB_A, K, 0, NBMAX, SCALED>;
    auto kern_b = sos_tma_kernel<PFB_T, PFB_A, K, 1, NBMAX, SCALED>;
    auto kern_l = sos_tma_kernel<PFB_T, PFB_A, K, 2, NBMAX, SCALED>;
    const int smem_a = kTmaStages * 2 * kTileBytes + 3 * kTmaStages * 8;
    const int smem_b = kTmaStages * 2 * kTileBytes + p.nb * out_tile_bytes<PFB_T, PFB_A>(2, false, false) + 3 * kTmaStages * 8;
    const int smem_bd = kTmaStages * 2 * kTileBytes + p.nb * out_tile_bytes<PFB_T, PFB_A>(2, true, false) + 3 * kTmaStages * 8;
    cudaError_t e = cudaFuncSetAttribute(kern_a, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern_b, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern_l, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);
    if (e != cudaSuccess) return (int)e;
    if (p.weighted) return (int)cudaErrorInvalidValue;   // run_tma_weighted
    if (p.dec_band1 && (!PFB_HAS_DEC || p.levels || p.wide || p.weighted)) return (int)cudaErrorInvalidValue;
    static thread_local CoefTable<PFB_A> tab;
    fill_tma_table<K, SCALED>(p, coef_host, tab);
    constexpr int S = 2 * K;
    constexpr int LPC = S <= 2 ? 2 : S <= 4 ? 4 : S <= 8 ? 8 : 16;
    const unsigned grid = (unsigned)p.CB * (unsigned)p.G * (unsigned)p.NW;
    const unsigned grid_a = (unsigned)p.CB * (unsigned)p.G_a * (unsigned)p.NW;
    const int threads = (p.nb + 1) * 32;
    if ((phases & 1) && p.J == 1 && ev) {   // a single chunk per cascade needs no pre-pass
        cudaEventRecord(ev[0], s);
        cudaEventRecord(ev[1], s);
    }
    if ((phases & 1) && p.J > 1) {
        if (ev) cudaEventRecord(ev[0], s);
        if (p.H) {   // GEMM-form pre-pass (float64 DMMA for every arithmetic type) when the bank has its tables
            auto kern_h = sos_hform_kernel<PFB_T, PFB_A, K, NBMAX>;
            e = cudaFuncSetAttribute(kern_h, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);
            if (e != cudaSuccess) return (int)e;
            kern_h<<<grid_a, (p.nb_a + 1) * 32, smem_a, s>>>(p, xmap);
        } else
            kern_a<<<grid_a, (p.nb_a + 1) * 32, smem_a, s>>>(p, tab, xmap, ymap, ymap2);
        if (ev) cudaEventRecord(ev[1], s);
        *launches += 1;
    }
    if (phases & 2) {
        if (p.J > 1) {
            sos_carry_kernel<PFB_A, K><<<(unsigned)(((long long)p.Q * LPC + 127) / 128), 128, 0, s>>>(p);
            *launches += 1;
        }
        if (ev) cudaEventRecord(ev[2], s);
        if (p.dec_band1) {
#if PFB_HAS_DEC
            auto kern_bd = sos_tma_kernel<PFB_T, PFB_A, K, 1, NBMAX, SCALED, false, true>;
            e = cudaFuncSetAttribute(kern_bd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bd);
            if (e != cudaSuccess) return (int)e;
            kern_bd<<<grid, threads, smem_bd, s>>>(p, tab, xmap, ymap, ymap2);
#endif
        } else if (p.levels) {
            kern_l<<<grid, threads, smem_a, s>>>(p, tab, xmap, ymap, ymap2);   // energies instead of band signals
        } else if (p.wide) {
            auto kern_w = sos_tma_wide_kernel<PFB_T, PFB_A, K, SCALED>;
            const int smem_w = 2 * 4 * kTileBytes + p.nb * out_tile_bytes<PFB_T, PFB_A>(4, false, false) + 3 * 2 * 8;
            e = cudaFuncSetAttribute(kern_w, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_w);
            if (e != cudaSuccess) return (int)e;
            kern_w<<<grid, threads, smem_w, s>>>(p, tab, xmap, ymap);
        } else {
            kern_b<<<grid, threads, smem_b, s>>>(p, tab, xmap, ymap, ymap2);
        }
        if (ev) cudaEventRecord(ev[3], s);
        *launches += 1;
    }
    return (int)cudaGetLastError();
}
template <int K, int NBMAX>
static int run_tma_n(const TmaParams &p, const double *coef_host, const CUtensorMap &xmap, const CUtensorMap &ymap,
                     const CUtensorMap &ymap2, cudaStream_t s, int *launches, cudaEvent_t *ev, int phases) {
#if PFB_HAS_SCALED
    if (p.gains) return run_tma_ns<K, NBMAX, true>(p, coef_host, xmap, ymap, ymap2, s, launches, ev, phases);
#endif
    return run_tma_ns<K, NBMAX, false>(p, coef_host, xmap, ymap, ymap2, s, launches, ev, phases);
}
template <int K>
static int run_tma(const TmaParams &p, const double *coef_host, const CUtensorMap &xmap, const CUtensorMap &ymap,
                   const CUtensorMap &ymap2, cudaStream_t s, int *launches, cudaEvent_t *ev, int phases) {
    if (p.weighted) {
#if PFB_HAS_SCALED
        if (p.gains) return run_tma_weighted<K, true>(p, coef_host, xmap, ymap, s, launches, ev, phases);
#endif
        return run_tma_weighted<K, false>(p, coef_host, xmap, ymap, s, launches, ev, phases);
    }
    return p.nb <= kTmaSmallBands ? run_tma_n<K, kTmaSmallBands>(p, coef_host, xmap, ymap, ymap2, s, launches, ev, phases)
                                  : run_tma_n<K, kTmaMaxBands>(p, coef_host, xmap, ymap, ymap2, s, launches, ev, phases);
}

#define PFB_K_SWITCH(CALL)                          \
    switch (K) {                                    \
        case 1: return CALL(1);                     \
        case 2: return CALL(2);                     \
        case 3: return CALL(3);                     \
        case 4: return CALL(4);                     \
        case 6: return CALL(6);                     \
        case 8: return CALL(8);                     \
        default: return (int)cudaErrorInvalidValue; \
    }

int PFB_CAT(launch_seq_, PFB_SUFFIX)(const SosParams &p, int K, bool exact, cudaStream_t s) {
#define PFB_SEQ_FAST(k) run_seq<k, false>(p, s)
#define PFB_SEQ_EXACT(k) run_seq<k, PFB_HAS_EXACT>(p, s)
    if (exact) {
        PFB_K_SWITCH(PFB_SEQ_EXACT)
    }
    PFB_K_SWITCH(PFB_SEQ_FAST)
}

int PFB_CAT(launch_scan_, PFB_SUFFIX)(SosParams p, int K, int warps, const double *coef_host, int rows, int C,
                                      cudaStream_t s, int *launches) {
#define PFB_SCAN(k) run_scan<k>(p, warps, coef_host, rows, C, s, launches)
    PFB_K_SWITCH(PFB_SCAN)
}

int PFB_CAT(launch_tma_, PFB_SUFFIX)(const TmaParams &p, int K, const double *coef_host, const CUtensorMap &xmap,
                                     const CUtensorMap &ymap, const CUtensorMap &ymap2, cudaStream_t s, int *launches,
                                     cudaEvent_t *ev, int phases) {
#define PFB_TMA(k) run_tma<k>(p, coef_host, xmap, ymap, ymap2, s, launches, ev, phases)
    PFB_K_SWITCH(PFB_TMA)
}

}  // namespace pfb
