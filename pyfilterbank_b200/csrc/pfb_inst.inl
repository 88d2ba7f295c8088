// pfb_inst.inl -- kernel instantiations and launchers for one (signal type, arithmetic type) pair.
// Included by pfb_inst_{dd,ff,fd}.cu with PFB_T, PFB_A and PFB_SUFFIX defined.
#include <vector>

#include "pfb_kernels.cuh"

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

}  // namespace pfb
