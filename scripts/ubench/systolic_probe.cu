// Microbenchmark (compute only): two ways to run a 4-section normalised biquad cascade.
//   lane_probe : one lane per cascade, coefficients uniform per warp (kernel parameters -> uniform registers),
//                17 FP64 ops per band-sample (the scheme of sos_tma_kernel)
//   sys_probe  : one lane per SECTION (4 lanes per cascade), each section hands its output to the next lane
//                with a shuffle; coefficients per lane in registers; 5 FP64 ops per lane-sample = 20 per band-sample
// Reports band-samples/s for the whole GPU at several occupancies.
#include <cstdio>
#include <cuda_runtime.h>

struct Coef { double c[4][5]; };

__global__ void lane_probe(double *out, const __grid_constant__ Coef cf, int iters) {
    double w[4][2] = {};
    double x = threadIdx.x * 1e-3, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int u = 0; u < 8; ++u) {
            double s = x;
            x = x * 0.999 + 1e-3;   // next input sample (stands for the shared-memory read)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double t = fma(-cf.c[k][4], w[k][1], s);
                const double w0 = fma(-cf.c[k][3], w[k][0], t);
                s = fma(cf.c[k][2], w[k][1], fma(cf.c[k][1], w[k][0], w0));
                w[k][1] = w[k][0];
                w[k][0] = w0;
            }
            acc += s * cf.c[0][0];
        }
    }
    if (acc == 123.456) out[0] = acc;
}

__global__ void sys_probe(double *out, const double *coef, int iters) {
    const int lane = threadIdx.x & 31, sec = lane & 3;
    const double g = coef[sec * 5 + 0], c1 = coef[sec * 5 + 1], c2 = coef[sec * 5 + 2], a1 = coef[sec * 5 + 3],
                 a2 = coef[sec * 5 + 4];
    double w1 = 0, w2 = 0, y = 0, acc = 0;
    double x = threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int u = 0; u < 8; ++u) {
            double xin = __shfl_up_sync(0xffffffffu, y, 1);
            x = x * 0.999 + 1e-3;
            if (sec == 0) xin = x;
            const double t = fma(-a2, w2, xin);
            const double w0 = fma(-a1, w1, t);
            y = fma(c2, w2, fma(c1, w1, g * w0));
            w2 = w1;
            w1 = w0;
            acc += y;
        }
    }
    if (acc == 123.456) out[0] = acc;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out, *coef; cudaMalloc(&out, 8); cudaMalloc(&coef, 20 * 8);
    double hc[20]; Coef cf;
    for (int k = 0; k < 4; ++k) { double v[5] = {1.0, 2.0, 1.0, -1.9, 0.95}; for (int i = 0; i < 5; ++i) { hc[k * 5 + i] = v[i]; cf.c[k][i] = v[i]; } }
    cudaMemcpy(coef, hc, sizeof hc, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    for (int warps : {8, 12, 16, 24, 32}) {
        float ms;
        lane_probe<<<sms, warps * 32>>>(out, cf, 10);
        cudaEventRecord(e0); lane_probe<<<sms, warps * 32>>>(out, cf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double bs_lane = (double)sms * warps * 32 * iters * 8 / (ms * 1e-3);
        sys_probe<<<sms, warps * 32>>>(out, coef, 10);
        cudaEventRecord(e0); sys_probe<<<sms, warps * 32>>>(out, coef, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms2; cudaEventElapsedTime(&ms2, e0, e1);
        const double bs_sys = (double)sms * warps * 8 * iters * 8 / (ms2 * 1e-3);
        printf("warps/SM=%2d  lane-per-cascade: %.3e band-samples/s (%.1f DFMA/clk/SM @1.9GHz)   section-per-lane: %.3e band-samples/s (%.1f FP64 ops/clk/SM)\n",
               warps, bs_lane, bs_lane * 18 / sms / 1.9e9, bs_sys, bs_sys * 21 / sms / 1.9e9);
    }
    return 0;
}
