// Microbenchmark: do FP64 tensor-core MMAs (DMMA) and scalar DFMAs share an issue pipe on this GPU?
// Runs DFMA-only, DMMA-only and interleaved kernels with identical per-warp instruction counts; if
// t(mixed) ~ max(t_dfma, t_dmma) the pipes are independent, if ~ sum they are shared.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: DFMA only (NF per iteration), 1: DMMA m8n8k4 only (NM per iteration), 2: both, 3: m16n8k16 only, 4: m16n8k16 + DFMA
template <int MODE, int NF, int NM>
__global__ void k(double *out, double a, double b, int iters) {
    double v[8], acc[8][2], acc4[4][4];
    double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x + i; acc[i][0] = acc[i][1] = 0; av[i] = a + i; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { bv[i] = b + i; for (int j = 0; j < 4; ++j) acc4[i][j] = 0; }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 4) {
#pragma unroll
            for (int r = 0; r < NF; ++r) v[r & 7] = fma(v[r & 7], a, b);
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int r = 0; r < NM; ++r) dmma884(acc[r & 7][0], acc[r & 7][1], a, b);
        }
        if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int r = 0; r < NM; ++r) dmma16816(acc4[r & 3], av, bv);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i] + acc[i][0] + acc[i][1];
#pragma unroll
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc4[i][j];
    if (s == 123.456) out[0] = s;
}

template <int MODE, int NF, int NM>
float run(const char *name, int warps, int sms, int iters) {
    double *out; cudaMalloc(&out, 8);
    k<MODE, NF, NM><<<sms, warps * 32>>>(out, 1.0000001, 1e-9, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, NF, NM><<<sms, warps * 32>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (MODE == 0 || MODE == 2 || MODE == 4) ? (double)iters * NF * warps * sms : 0;   // warp instructions
    const double mmas = (MODE != 0) ? (double)iters * NM * warps * sms : 0;
    const double flop_per_mma = (MODE >= 3) ? 2.0 * 16 * 8 * 16 : 2.0 * 8 * 8 * 4;
    printf("%-28s warps/SM=%2d: %8.3f ms | DFMA %.2f warp-inst/clk/SM (%.1f TF) | DMMA %.3f inst/clk/SM (%.1f TF)\n", name, warps, ms,
           fmas / (ms * 1e-3) / sms / 1.9e9, fmas * 64 / (ms * 1e-3) / 1e12, mmas / (ms * 1e-3) / sms / 1.9e9,
           mmas * flop_per_mma / (ms * 1e-3) / 1e12);
    cudaFree(out);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s SMs=%d\n", p.name, sms);
    for (int warps : {8, 16, 32}) {
        run<0, 16, 0>("DFMA x16", warps, sms, 20000);
        run<1, 0, 8>("DMMA884 x8", warps, sms, 20000);
        run<2, 16, 8>("DFMA x16 + DMMA884 x8", warps, sms, 20000);
        run<1, 0, 2>("DMMA884 x2", warps, sms, 20000);
        run<2, 16, 2>("DFMA x16 + DMMA884 x2", warps, sms, 20000);
        run<3, 0, 4>("DMMA16816 x4", warps, sms, 5000);
        run<4, 16, 1>("DFMA x16 + DMMA16816 x1", warps, sms, 5000);
        run<3, 0, 1>("DMMA16816 x1", warps, sms, 5000);
    }
    return 0;
}
