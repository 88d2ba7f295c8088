// Microbenchmark: HBM write bandwidth of the output pass's store pattern.  Every warp owns 32 streams
// (time chunks, `L` doubles apart) and stores [32 streams x FRAG bytes] per step with TMA
// (cp.async.bulk.tensor.3d, 128-byte swizzled boxes of 32 rows x 128 B, FRAG/128 boxes per step),
// no arithmetic.  Also: plain contiguous streaming stores (pure-write ceiling) and a device copy.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BOXES>
__global__ void __launch_bounds__(256) tma_wr(const __grid_constant__ CUtensorMap ymap, long long L, int groups, int steps_per_wait) {
    extern __shared__ __align__(1024) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    char *buf = smem + warp * BOXES * 4096;
    for (int i = lane; i < BOXES * 512; i += 32) reinterpret_cast<double *>(buf)[i] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
    // cascade q = blockIdx.x / groups * 8 + warp ; chunk group w = blockIdx.x % groups
    const int q = (blockIdx.x / groups) * 8 + warp, w = blockIdx.x % groups;
    if (lane == 0) {
        const long long nt = L / (16 * BOXES);
        for (long long t = 0; t < nt; ++t) {
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            for (int b = 0; b < BOXES; ++b) {
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(&ymap),
                             "r"(smem_addr(buf + b * 4096)), "r"((int)(t * 16 * BOXES + 16 * b)), "r"(32 * w), "r"(q)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
}

// The same store pattern with ONE-DIMENSIONAL bulk copies: every lane copies its own row fragment (FRAG contiguous
// bytes, shared -> global) with cp.async.bulk, all 32 lanes in one instruction; rows are FRAG + 16 bytes apart in shared
// memory (what a lane-per-row writer needs to stay free of bank conflicts).  32 requests of FRAG bytes per step instead
// of FRAG / 128 tensor requests of 32 x 128 bytes issued by one lane.
template <int FRAG>
__global__ void __launch_bounds__(256) bulk1d_wr(double *y, long long L, long long n, int groups) {
    extern __shared__ __align__(1024) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int RS = FRAG + 16;
    char *buf = smem + warp * 32 * RS;
    for (int i = lane; i < 32 * RS / 8; i += 32) reinterpret_cast<double *>(buf)[i] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
    const int q = (blockIdx.x / groups) * 8 + warp, w = blockIdx.x % groups;
    char *dst = reinterpret_cast<char *>(y + (size_t)q * n + (size_t)(32 * w + lane) * L);
    const uint32_t src = smem_addr(buf + lane * RS);
    const long long nt = L * 8 / FRAG;
    for (long long t = 0; t < nt; ++t) {
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst + t * FRAG), "r"(src), "r"(FRAG) : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

// The same store pattern with plain global stores issued COOPERATIVELY: per step the warp writes its 32 row fragments one
// after the other, every instruction one contiguous piece of 512 bytes (32 lanes x 16 bytes) of ONE row -- no TMA.
// CS: 0 default stores, 1 st.global.cs (streaming)
template <int FRAG, int CS>
__global__ void __launch_bounds__(256) coop_wr(double *y, long long L, long long n, int groups) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = (blockIdx.x / groups) * 8 + warp, w = blockIdx.x % groups;
    char *base = reinterpret_cast<char *>(y + (size_t)q * n + (size_t)(32 * w) * L);     // row r: base + r * L * 8
    const long long nt = L * 8 / FRAG;
    const double v0 = 1.0 + lane, v1 = 2.0;
    for (long long t = 0; t < nt; ++t) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            char *row = base + (size_t)r * L * 8 + t * FRAG;
#pragma unroll
            for (int c = 0; c < FRAG / 512; ++c) {
                double *p = reinterpret_cast<double *>(row + c * 512 + lane * 16);
                if (CS) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v0), "d"(v1) : "memory");
                else asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v0), "d"(v1) : "memory");
            }
        }
    }
}

// HINT: 1 evict_first, 2 evict_last, 3 evict_unchanged (createpolicy.fractional), BOXES boxes per step
template <int BOXES, int HINT>
__global__ void __launch_bounds__(256) tma_wr_hint(const __grid_constant__ CUtensorMap ymap, long long L, int groups, int) {
    extern __shared__ __align__(1024) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    char *buf = smem + warp * BOXES * 4096;
    for (int i = lane; i < BOXES * 512; i += 32) reinterpret_cast<double *>(buf)[i] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
    const int q = (blockIdx.x / groups) * 8 + warp, w = blockIdx.x % groups;
    uint64_t pol;
    if (HINT == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    else if (HINT == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;\n" : "=l"(pol));
    if (lane == 0) {
        const long long nt = L / (16 * BOXES);
        for (long long t = 0; t < nt; ++t) {
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            for (int b = 0; b < BOXES; ++b) {
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;\n" ::"l"(&ymap),
                             "r"(smem_addr(buf + b * 4096)), "r"((int)(t * 16 * BOXES + 16 * b)), "r"(32 * w), "r"(q), "l"(pol)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
}

// every lane streams its own chunk with VB-byte vector stores, FRAG bytes per step (then a pause of `spin` clocks)
template <int VB>
__global__ void __launch_bounds__(256) lane_wr(double *y, long long L, long long n, int groups, int frag_el) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = (blockIdx.x / groups) * 8 + warp, w = blockIdx.x % groups;
    double *p = y + (size_t)q * n + (size_t)(32 * w + lane) * L;
    for (long long e = 0; e < L; e += frag_el) {
        if (VB == 32) {
#pragma unroll 4
            for (int i = 0; i < frag_el; i += 4)
                asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p + e + i), "d"(1.0), "d"(2.0), "d"(3.0), "d"(4.0) : "memory");
        } else {
#pragma unroll 4
            for (int i = 0; i < frag_el; i += 2)
                asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p + e + i), "d"(1.0), "d"(2.0) : "memory");
        }
    }
}

__global__ void contiguous_wr(double2 *y, size_t n2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(y + i), "d"(1.0), "d"(2.0) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int sweep_L(EncodeTiledFn enc);

int main(int argc, char **argv) {
    void *f = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
    EncodeTiledFn enc = (EncodeTiledFn)f;
    if (argc > 1) return sweep_L(enc);
    const int Q = 2112;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int groups : {2, 6}) {
        const int J = 32 * groups;
        const long long L = 2880000 / J / 128 * 128, n = L * J;
        double *y; const size_t bytes = (size_t)Q * n * 8;
        if (cudaMalloc(&y, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
        CUtensorMap map;
        cuuint64_t dims[3] = {(cuuint64_t)L, (cuuint64_t)J, (cuuint64_t)Q};
        cuuint64_t strides[2] = {(cuuint64_t)L * 8, (cuuint64_t)n * 8};
        cuuint32_t box[3] = {16, 32, 1}, estr[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        auto run = [&](auto kern, int boxes) {
            const int smem = 8 * boxes * 4096;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            float best = 1e9;
            for (int it = 0; it < 3; ++it) {
                cudaEventRecord(e0);
                kern<<<Q / 8 * groups, 256, smem>>>(map, L, groups, 1);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("TMA store  J=%3d  fragment %4d B: %7.3f ms  %.0f GB/s  (%s)\n", J, boxes * 128, best, bytes / (best * 1e-3) / 1e9,
                   cudaGetErrorString(cudaGetLastError()));
        };
        run(tma_wr<1>, 1); run(tma_wr<2>, 2); run(tma_wr<4>, 4);
        printf("evict_first:     "); run(tma_wr_hint<2, 1>, 2);
        printf("evict_last:      "); run(tma_wr_hint<2, 2>, 2);
        printf("evict_unchanged: "); run(tma_wr_hint<2, 3>, 2);
        printf("evict_last 512:  "); run(tma_wr_hint<4, 2>, 4);
        auto run1d = [&](auto kern, int frag) {
            const int smem = 8 * 32 * (frag + 16);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            float best = 1e9;
            for (int it = 0; it < 3; ++it) {
                cudaEventRecord(e0);
                kern<<<Q / 8 * groups, 256, smem>>>(y, L, n, groups);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("1-D bulk copy per lane  J=%3d  fragment %4d B: %7.3f ms  %.0f GB/s  (%s)\n", J, frag, best, bytes / (best * 1e-3) / 1e9,
                   cudaGetErrorString(cudaGetLastError()));
        };
        run1d(bulk1d_wr<128>, 128); run1d(bulk1d_wr<256>, 256); run1d(bulk1d_wr<512>, 512); run1d(bulk1d_wr<1024>, 1024);
        auto runc = [&](auto kern, const char *name) {
            float best = 1e9;
            for (int it = 0; it < 3; ++it) {
                cudaEventRecord(e0);
                kern<<<Q / 8 * groups, 256>>>(y, L, n, groups);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("cooperative global stores, %s  J=%3d: %7.3f ms  %.0f GB/s  (%s)\n", name, J, best, bytes / (best * 1e-3) / 1e9,
                   cudaGetErrorString(cudaGetLastError()));
        };
        runc(coop_wr<512, 0>, "512-byte row pieces, st.global   "); runc(coop_wr<512, 1>, "512-byte row pieces, st.global.cs");
        runc(coop_wr<1024, 1>, "1024-byte row pieces, st.global.cs");
        for (int vb : {16, 32}) {
            float best = 1e9;
            for (int it = 0; it < 3; ++it) {
                cudaEventRecord(e0);
                if (vb == 32) lane_wr<32><<<Q / 8 * groups, 256>>>(y, L, n, groups, 32);
                else lane_wr<16><<<Q / 8 * groups, 256>>>(y, L, n, groups, 32);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("per-lane st.global.cs %d B vectors: %7.3f ms  %.0f GB/s (%s)\n", vb, best, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
        float best = 1e9;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            contiguous_wr<<<148 * 16, 256>>>((double2 *)y, bytes / 16);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("contiguous st.cs.v2.f64 (write only): %7.3f ms  %.0f GB/s\n", best, bytes / (best * 1e-3) / 1e9);
        best = 1e9;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            cudaMemsetAsync(y, 0, bytes);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("cudaMemset: %7.3f ms  %.0f GB/s\n", best, bytes / (best * 1e-3) / 1e9);
        best = 1e9;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            cudaMemcpyAsync(y, (char *)y + bytes / 2, bytes / 2, cudaMemcpyDeviceToDevice);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("cudaMemcpy D2D (read+write bytes): %7.3f ms  %.0f GB/s\n", best, bytes / (best * 1e-3) / 1e9);
        cudaFree(y);
    }
    return 0;
}

// chunk-length sensitivity of the 256-byte-fragment store pattern: J = 256 chunks, row stride = L doubles
int sweep_L(EncodeTiledFn enc) {
    const int Q = 2112, J = 256, groups = 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long Lmax = 12544;
    double *y; const size_t cap = (size_t)Q * Lmax * J * 8;
    if (cudaMalloc(&y, cap) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    for (long long L : {11232LL, 11264LL, 11200LL, 11136LL, 10240LL, 12288LL, 8192LL, 11232LL + 32, 11232LL - 32, 9984LL, 12000LL, 12544LL, 4096LL, 5632LL}) {
        const long long n = L * J;
        const size_t bytes = (size_t)Q * n * 8;
        CUtensorMap map;
        cuuint64_t dims[3] = {(cuuint64_t)L, (cuuint64_t)J, (cuuint64_t)Q};
        cuuint64_t strides[2] = {(cuuint64_t)L * 8, (cuuint64_t)n * 8};
        cuuint32_t box[3] = {16, 32, 1}, estr[3] = {1, 1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
        const int smem = 8 * 2 * 4096;
        cudaFuncSetAttribute(tma_wr<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        float best = 1e9;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            tma_wr<2><<<Q / 8 * groups, 256, smem>>>(map, L, groups, 1);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("J=256 256-B fragments, chunk length L=%6lld (row stride %8lld B): %7.3f ms  %.0f GB/s\n", L, L * 8, best, bytes / (best * 1e-3) / 1e9);
    }
    return 0;
}
