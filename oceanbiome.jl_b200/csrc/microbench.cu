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

// ---- many-stream access pattern ---------------------------------------------------------------------
// The memory system's ceiling for the access pattern of a fused tendency kernel: every thread reads its
// cell from `nread` separate fields and read-modify-writes (mode 0) or writes (mode 1) `nrmw` fields,
// same launch geometry as the real kernels, no arithmetic to speak of.  A plain two-stream copy (the
// MEASURED_PEAKS.json number) does not capture what 40–60 concurrent streams do to DRAM efficiency.
namespace obm {
constexpr int SP_MAX_READ = 40, SP_MAX_RMW = 26;
struct StreamArgs {
    GridDims d;
    const double* rd[SP_MAX_READ];
    double* wr[SP_MAX_RMW];
    int nread, nrmw, mode;
};
__global__ void __launch_bounds__(128) stream_pattern_kernel(const __grid_constant__ StreamArgs a) {
    int i, j, k;
    if (!thread_cell(a.d, i, j, k)) return;
    const long long idx = cell_index(a.d, i, j, k);
    double v[SP_MAX_READ];
#pragma unroll
    for (int n = 0; n < SP_MAX_READ; n++) v[n] = n < a.nread ? a.rd[n][idx] : 0.0;
    double s = 0;
#pragma unroll
    for (int n = 0; n < SP_MAX_READ; n++) s += v[n];
    if (a.mode == 0) {
        double o[SP_MAX_RMW];
#pragma unroll
        for (int n = 0; n < SP_MAX_RMW; n++) o[n] = n < a.nrmw ? a.wr[n][idx] : 0.0;
#pragma unroll
        for (int n = 0; n < SP_MAX_RMW; n++)
            if (n < a.nrmw) a.wr[n][idx] = o[n] + s * 1e-300;
    } else {
#pragma unroll
        for (int n = 0; n < SP_MAX_RMW; n++)
            if (n < a.nrmw) a.wr[n][idx] = s;
    }
}
}  // namespace obm

// Diagnostic (synchronises): GB/s moved (8·cells·(nread + (mode == 0 ? 2 : 1)·nrmw) per launch) over `reps`
// launches; negative on error.
extern "C" double obm_stream_pattern_gbs(const obm_grid* grid, int nread, const double* const* reads, int nrmw,
                                         double* const* rmw, int mode, int reps, void* stream) {
    using namespace obm;
    if (!reads || !rmw || nread < 0 || nread > SP_MAX_READ || nrmw < 0 || nrmw > SP_MAX_RMW || reps <= 0)
        return (double)OBM_ESIZE;
    static thread_local StreamArgs a;
    if (make_dims(grid, &a.d, false)) return (double)OBM_ESIZE;
    for (int n = 0; n < nread; n++) a.rd[n] = reads[n];
    for (int n = 0; n < nrmw; n++) a.wr[n] = rmw[n];
    a.nread = nread; a.nrmw = nrmw; a.mode = mode;
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    stream_pattern_kernel<<<cell_grid(a.d, 128), 128, 0, s>>>(a);
    cudaEventRecord(e0, s);
    for (int r = 0; r < reps; r++) stream_pattern_kernel<<<cell_grid(a.d, 128), 128, 0, s>>>(a);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    int rc = launch_status("stream_pattern_kernel");
    if (rc) return -(double)rc;
    const double bytes = 8.0 * (double)cell_count(a.d) * (nread + (mode == 0 ? 2 : 1) * nrmw) * reps;
    return bytes / (ms * 1e-3) / 1e9;
}

// ---- instruction-fetch ceiling ------------------------------------------------------------------------
// The PISCES tendency kernel is ≈ 3 200 straight-line instructions that every warp executes once.  This pair of kernels
// does the same FP64 work (3 072 DFMA per thread, 8 independent chains) once as straight-line code and once as a
// 48-instruction loop body: the ratio of their run times is what a single pass over a long instruction stream costs
// on this device, independent of memory traffic and register pressure.
namespace obm {
#define OBM_R8(x) x x x x x x x x
#define OBM_STEP                                                                                  \
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);               \
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
template <bool STRAIGHT>
__global__ void __launch_bounds__(128) fetch_kernel(double* out, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    if (STRAIGHT) {
        OBM_R8(OBM_R8(OBM_R8(OBM_STEP)) ) /* 512 steps … */
        x0 += 0;
    } else {
#pragma unroll 1
        for (int i = 0; i < 64; i++) { OBM_R8(OBM_STEP) }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
#undef OBM_STEP
#undef OBM_R8
}  // namespace obm

// Diagnostic (synchronises): ms for `blocks` blocks of 128 threads, each thread 4 096 DFMA, as straight-line code
// (straight != 0) or as a loop.  scratch: DEVICE buffer of >= blocks*128 doubles.
extern "C" double obm_fetch_ceiling_ms(double* scratch, int blocks, int straight, void* stream) {
    using namespace obm;
    if (!scratch || blocks <= 0) return (double)OBM_ENULL;
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        if (rep == 1) cudaEventRecord(e0, s);
        if (straight) fetch_kernel<true><<<blocks, 128, 0, s>>>(scratch, 0.999999, 1e-9);
        else fetch_kernel<false><<<blocks, 128, 0, s>>>(scratch, 0.999999, 1e-9);
    }
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    int rc = launch_status("fetch_kernel");
    if (rc) return -(double)rc;
    return (double)ms;
}
