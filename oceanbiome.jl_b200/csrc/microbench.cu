// microbench.cu — FP64 pipe peak (DFMA/s), measured on the device this library runs on.  The FP64
// roofline of the transcendental-heavy kernels (PAR, carbonate, PISCES) is quoted against this number
// (SURVEY §8d asks for a measured DFMA denominator next to the measured HBM bandwidth).
#include "obm_common.cuh"

namespace obm {
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace obm

// Diagnostic (synchronises!): returns DFMA instructions per second (per thread-op, i.e. 2 FLOP each),
// or a negative error.  `scratch` = device buffer of at least 148*8*256 doubles.
extern "C" double obm_fp64_peak_dfma_per_s(double* scratch, int iters, void* stream) {
    if (!scratch || iters <= 0) return (double)OBM_ENULL;
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    obm::dfma_kernel<<<blocks, 256, 0, s>>>(scratch, 64, 0.999999, 1e-9);  // warm-up
    cudaEventRecord(e0, s);
    obm::dfma_kernel<<<blocks, 256, 0, s>>>(scratch, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    int rc = obm::launch_status("dfma_kernel");
    if (rc) return -(double)rc;
    return (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3);
}
