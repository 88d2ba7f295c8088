// Microbenchmark: HBM write / read bandwidth when every lane of every warp owns its own stream
// (streams `stride` bytes apart) and a warp moves tiles of [32 streams x R bytes] per step.
// This is the access pattern of the time-chunked SOS kernel; R is the tile row length.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void wr(double *y, long long stride_el, long long len_el, int R_el, int mode) {
    // warp w handles streams [32w, 32w+32); stream s starts at y + s*stride_el, len_el elements each
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double *base = y + warp * 32 * stride_el;
    const int segs = R_el / 2;              // 16-byte segments per row
    double2 v = make_double2(1.0, 2.0);
    double2 acc = make_double2(0, 0);
    for (long long e0 = 0; e0 < len_el; e0 += R_el) {
        // 32 rows x segs segments, lanes sweep segments of a row first (coalesced)
        for (int i = lane; i < 32 * segs; i += 32) {
            const int row = i / segs, seg = i % segs;
            double2 *p = reinterpret_cast<double2 *>(base + row * stride_el + e0 + seg * 2);
            if (mode == 0) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
            else { double2 t; asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "l"(p)); acc.x += t.x; acc.y += t.y; }
        }
    }
    if (mode && acc.x == 123.25) y[0] = acc.y;
}
int main(int argc, char **argv) {
    const long long nstreams = 2112LL * 32 * 2;          // 135168 streams like cfg 2 with 2 warps per cascade
    const long long len_el = 45056;                      // elements per stream (352 KB)
    const long long stride_el = len_el;
    double *y; size_t bytes = nstreams * stride_el * 8;
    if (cudaMalloc(&y, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(y, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode)
        for (int wpb : {2, 8})
            for (int R_el : {16, 32, 64, 128, 256, 512}) {
                const int threads = wpb * 32; const int blocks = (int)(nstreams / 32 / wpb);
                float best = 1e9;
                for (int it = 0; it < 3; ++it) {
                    cudaEventRecord(e0);
                    wr<<<blocks, threads>>>(y, stride_el, len_el, R_el, mode);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
                }
                printf("%s warps/block=%d row=%4d B: %.2f ms  %.0f GB/s\n", mode ? "read " : "write", wpb, R_el * 8, best, bytes / (best * 1e-3) / 1e9);
            }
    return 0;
}
