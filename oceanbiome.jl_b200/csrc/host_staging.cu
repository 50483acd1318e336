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
