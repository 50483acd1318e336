// host_staging.cu — host-buffer staging for callers whose fields live in HOST memory (the end-to-end path
// of bench.py): strided slab copies of halo'd parent arrays between pinned host memory and the device,
// so that a stage can be pipelined slab by slab (H2D of slab s+1 ∥ kernels of slab s ∥ D2H of slab s−1).
// Every hot kernel is pointwise or column-local, so x–y slabs are independent.
#include "obm_common.cuh"

using namespace obm;

// Copies interior rows j ∈ [j0, j1) (all x incl. halos) of `nplanes` k-planes of each field.
// direction: 0 = host → device, 1 = device → host.  One cudaMemcpy2DAsync per field:
// width = (j1 − j0)·(Nx + 2Hx) doubles, pitch = one x–y plane.
extern "C" int obm_copy_slab(const obm_grid* grid, int nfields, void* const* dst, const void* const* src, int nplanes,
                             int direction, void* stream) {
    OBM_REQUIRE(dst && src, OBM_ENULL, "obm_copy_slab: dst / src is NULL");
    OBM_REQUIRE(nfields >= 0 && nplanes >= 1 && (direction == 0 || direction == 1), OBM_ESIZE,
                "obm_copy_slab: nfields = %d, nplanes = %d, direction = %d", nfields, nplanes, direction);
    GridDims d;
    int rc = make_dims(grid, &d, false);
    if (rc) return rc;
    const size_t row = (size_t)d.sy * sizeof(double);
    const size_t pitch = (size_t)d.sz * sizeof(double);
    const size_t width = (size_t)(d.j1 - d.j0) * row;
    const size_t offset = (size_t)(d.j0 + d.Hy) * row;
    const cudaMemcpyKind kind = direction == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    for (int f = 0; f < nfields; f++) {
        OBM_REQUIRE(dst[f] && src[f], OBM_ENULL, "obm_copy_slab: field %d is NULL", f);
        cudaError_t e = cudaMemcpy2DAsync((char*)dst[f] + offset, pitch, (const char*)src[f] + offset, pitch, width,
                                          (size_t)nplanes, kind, (cudaStream_t)stream);
        if (e != cudaSuccess) {
            set_error("obm_copy_slab: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    return 0;
}

// ---- the same slab copy driven by the SMs instead of the copy engines ------------------------------------------------
// cudaMemcpy2DAsync turns every (field, k-plane) row of a slab into its own DMA descriptor: with 32 slabs of a
// 1024-wide grid that is ≈ 10⁵ descriptors of 264 KB per direction per stage, and the per-descriptor cost shows
// (≈ 42 GB/s per direction against 49.6 GB/s for one large copy, scripts/pcie_bw.py).  Pinned host memory is mapped
// into the device's address space (UVA), so a small persistent kernel can stream the same rows with 16-byte
// loads/stores over PCIe: no descriptors, and one launch per slab.  One block per SM, 128 threads, 8 independent
// 16-byte accesses in flight per thread (2.4 MB in flight ≫ PCIe latency × bandwidth); the blocks are light enough
// (≈ 40 registers) to sit beside the compute kernels of the neighbouring slab.
namespace obm {

constexpr int COPY_MAX_FIELDS = 32;
constexpr int COPY_THREADS = 128;
constexpr int COPY_UNROLL = 8;

struct CopyArgs {
    char* dst[COPY_MAX_FIELDS];
    const char* src[COPY_MAX_FIELDS];
    int nfields, nplanes;
    size_t offset, pitch, width;  // bytes
};

template <typename V>
__global__ void __launch_bounds__(COPY_THREADS) copy_slab_kernel(const __grid_constant__ CopyArgs a) {
    const long long nv = (long long)(a.width / sizeof(V));
    const int units = a.nfields * a.nplanes;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int f = u / a.nplanes, p = u - f * a.nplanes;
        const V* __restrict__ s = reinterpret_cast<const V*>(a.src[f] + a.offset + (size_t)p * a.pitch);
        V* __restrict__ d = reinterpret_cast<V*>(a.dst[f] + a.offset + (size_t)p * a.pitch);
        for (long long v0 = threadIdx.x; v0 < nv; v0 += (long long)COPY_THREADS * COPY_UNROLL) {
            V r[COPY_UNROLL];
#pragma unroll
            for (int q = 0; q < COPY_UNROLL; q++) {
                const long long v = v0 + (long long)q * COPY_THREADS;
                if (v < nv) r[q] = s[v];
            }
#pragma unroll
            for (int q = 0; q < COPY_UNROLL; q++) {
                const long long v = v0 + (long long)q * COPY_THREADS;
                if (v < nv) d[v] = r[q];
            }
        }
    }
}

}  // namespace obm

// Same contract as obm_copy_slab; host pointers must be pinned (page-locked) memory, which UVA maps for the device.
extern "C" int obm_copy_slab_sm(const obm_grid* grid, int nfields, void* const* dst, const void* const* src, int nplanes,
                                int direction, void* stream) {
    OBM_REQUIRE(dst && src, OBM_ENULL, "obm_copy_slab_sm: dst / src is NULL");
    OBM_REQUIRE(nfields >= 0 && nplanes >= 1 && (direction == 0 || direction == 1), OBM_ESIZE,
                "obm_copy_slab_sm: nfields = %d, nplanes = %d, direction = %d", nfields, nplanes, direction);
    GridDims d;
    int rc = make_dims(grid, &d, false);
    if (rc) return rc;
    static thread_local CopyArgs a;
    const size_t row = (size_t)d.sy * sizeof(double);
    a.pitch = (size_t)d.sz * sizeof(double);
    a.width = (size_t)(d.j1 - d.j0) * row;
    a.offset = (size_t)(d.j0 + d.Hy) * row;
    a.nplanes = nplanes;
    for (int f0 = 0; f0 < nfields; f0 += COPY_MAX_FIELDS) {
        a.nfields = nfields - f0 < COPY_MAX_FIELDS ? nfields - f0 : COPY_MAX_FIELDS;
        bool wide = (a.pitch % 16 == 0) && (a.width % 16 == 0) && (a.offset % 16 == 0);
        for (int f = 0; f < a.nfields; f++) {
            OBM_REQUIRE(dst[f0 + f] && src[f0 + f], OBM_ENULL, "obm_copy_slab_sm: field %d is NULL", f0 + f);
            a.dst[f] = (char*)dst[f0 + f];
            a.src[f] = (const char*)src[f0 + f];
            wide = wide && ((uintptr_t)a.dst[f] % 16 == 0) && ((uintptr_t)a.src[f] % 16 == 0);
        }
        const int units = a.nfields * nplanes;
        const int blocks = units < 148 ? units : 148;
        if (wide) copy_slab_kernel<int4><<<blocks, COPY_THREADS, 0, (cudaStream_t)stream>>>(a);
        else copy_slab_kernel<double><<<blocks, COPY_THREADS, 0, (cudaStream_t)stream>>>(a);
        rc = launch_status("copy_slab_kernel");
        if (rc) return rc;
    }
    return 0;
}
