// pfb_tf.cuh -- launch interface of the transfer-function (scipy.signal.lfilter) kernel of pfb_fos_tf.cu, shared
// with pfb_api.cu (unfused remainder of the weighted bank call).
#pragma once
#include <cuda_runtime.h>

namespace pfb {

struct TfParams {
    const void *x;        // [R][ldx] float64 or float32
    void *y;              // [R][ldy], same type
    double *z;            // [R][max(ntaps - 1, 1)] in/out (scipy's zi convention)
    long long n, ldx, ldy;
    int R, ntaps;
    double b[8], a[8];    // a[0] == 1 (normalised on the host like scipy does)
};

// dtype PFB_F64 / PFB_F32 (float32 rows: float64 arithmetic, results rounded to float32); returns a cudaError_t
int launch_tf(const TfParams &p, int dtype, bool aligned, cudaStream_t s);

}  // namespace pfb
