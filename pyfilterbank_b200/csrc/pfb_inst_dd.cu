// pfb_inst_dd.cu -- kernel instantiations for signal type double, arithmetic type double.
#include "pfb_kernels.cuh"

namespace pfb {

template <int K, bool EXACT, bool ALIGNED>
static int run_seq_a(const SosParams &p, cudaStream_t s) {
    auto kern = sos_seq_kernel<double, double, K, EXACT, ALIGNED>;
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
static int run_scan_a(const SosParams &p, int warps, cudaStream_t s) {
    auto kern = sos_scan_kernel<double, double, K, ALIGNED>;
    const int smem = warps * 2 * kTileBytes + warps * 32 * 2 * K * (int)sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<(unsigned)p.Q, warps * 32, smem, s>>>(p);
    return (int)cudaGetLastError();
}

template <int K>
static int run_scan(const SosParams &p, int warps, cudaStream_t s) {
    return p.aligned ? run_scan_a<K, true>(p, warps, s) : run_scan_a<K, false>(p, warps, s);
}

#define PFB_K_SWITCH(CALL)            \
    switch (K) {                      \
        case 1: return CALL(1);       \
        case 2: return CALL(2);       \
        case 3: return CALL(3);       \
        case 4: return CALL(4);       \
        case 6: return CALL(6);       \
        case 8: return CALL(8);       \
        default: return (int)cudaErrorInvalidValue; \
    }

int launch_seq_dd(const SosParams &p, int K, bool exact, cudaStream_t s) {
#define SEQ_FAST(k) run_seq<k, false>(p, s)
#define SEQ_EXACT(k) run_seq<k, true>(p, s)
    if (exact) {
        PFB_K_SWITCH(SEQ_EXACT)
    }
    PFB_K_SWITCH(SEQ_FAST)
}

int launch_scan_dd(const SosParams &p, int K, int warps, cudaStream_t s) {
#define SCAN(k) run_scan<k>(p, warps, s)
    PFB_K_SWITCH(SCAN)
}

}  // namespace pfb
