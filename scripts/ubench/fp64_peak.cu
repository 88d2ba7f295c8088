// Microbenchmark: sustained DFMA / FFMA issue rate per SM on this GPU (roofline denominator for the
// compute-bound side of the SOS kernels).  ILP chains per thread, no memory traffic.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void fma_kernel(T *out, T a, T b, int iters) {
    T v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == (T)123.456) out[0] = s;
}
template <typename T, int ILP>
void run(const char *name, int warps_per_sm, int sms, double ghz_hint) {
    T *out; cudaMalloc(&out, 8);
    int iters = 20000;
    dim3 grid(sms), block(warps_per_sm * 32);
    fma_kernel<T, ILP><<<grid, block>>>(out, (T)1.0000001, (T)1e-9, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fma_kernel<T, ILP><<<grid, block>>>(out, (T)1.0000001, (T)1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)iters * 16 * ILP * warps_per_sm * 32 * sms;
    printf("%s ILP=%d warps/SM=%d: %.3f ms  %.3e FMA/s  => %.1f FMA/clk/SM at %.2f GHz, %.2f TFLOP/s\n", name, ILP,
           warps_per_sm, ms, fmas / (ms * 1e-3), fmas / (ms * 1e-3) / sms / (ghz_hint * 1e9), ghz_hint,
           2 * fmas / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    printf("%s SMs=%d clock=%.3f GHz\n", p.name, sms, ghz);
    run<double, 1>("f64", 4, sms, ghz);
    run<double, 1>("f64", 8, sms, ghz);
    run<double, 1>("f64", 16, sms, ghz);
    run<double, 2>("f64", 4, sms, ghz);
    run<double, 4>("f64", 4, sms, ghz);
    run<double, 8>("f64", 4, sms, ghz);
    run<double, 8>("f64", 16, sms, ghz);
    run<double, 8>("f64", 32, sms, ghz);
    run<float, 1>("f32", 4, sms, ghz);
    run<float, 1>("f32", 16, sms, ghz);
    run<float, 4>("f32", 4, sms, ghz);
    run<float, 8>("f32", 16, sms, ghz);
    run<float, 8>("f32", 32, sms, ghz);
    return 0;
}
