// timestepping.cu — the tracer update either side of the tendency pass, all tracers in ONE launch:
//     U[f] += Δt·(γ·Gⁿ[f] + ζ·G⁻[f]);   G⁻[f] ← Gⁿ[f]
// (rk3_substep! and cache_previous_tendencies! of src/BoxModel/timesteppers.jl:20-28,66-93; Oceananigans does the
// same per tracer for 3-D models: one launch per tracer per stage, then another per tracer for the cache copy).
// HBM-bound: 40 B per cell per tracer (read U, Gⁿ, G⁻; write U, G⁻), 24 B when there is no ζ term and no caching.
#include "obm_common.cuh"

namespace obm {

constexpr int TS_MAX_FIELDS = 40;
struct SubstepArgs {
    GridDims d;
    double* U[TS_MAX_FIELDS];
    const double* Gn[TS_MAX_FIELDS];
    double* Gm[TS_MAX_FIELDS];
    int nfields, has_zeta, cache;
    double dt, gamma, zeta;
};

__global__ void __launch_bounds__(256) rk3_substep_kernel(const __grid_constant__ SubstepArgs a) {
    int i, j, k;
    if (!thread_cell(a.d, i, j, k)) return;
    const long long idx = cell_index(a.d, i, j, k);
#pragma unroll 4
    for (int f = 0; f < a.nfields; f++) {
        const double gn = a.Gn[f][idx];
        double rhs;
        // convert(FT, Δt) * (γⁿ * Gⁿ + ζⁿ * G⁻), first stage: Δt * γ¹ * G¹ — no contraction, as Julia evaluates it
        if (a.has_zeta) rhs = __dmul_rn(a.dt, __dadd_rn(__dmul_rn(a.gamma, gn), __dmul_rn(a.zeta, a.Gm[f][idx])));
        else rhs = __dmul_rn(__dmul_rn(a.dt, a.gamma), gn);
        a.U[f][idx] = __dadd_rn(a.U[f][idx], rhs);
        if (a.cache) a.Gm[f][idx] = gn;
    }
}

}  // namespace obm

using namespace obm;

extern "C" int obm_rk3_substep(const obm_grid* grid, int nfields, double* const* U, const double* const* Gn,
                               double* const* Gm, double dt, double gamma, double zeta, int has_zeta,
                               int cache_previous, void* stream) {
    OBM_REQUIRE(nfields >= 0, OBM_ESIZE, "obm_rk3_substep: nfields = %d", nfields);
    if (nfields == 0) return 0;
    OBM_REQUIRE(U && Gn, OBM_ENULL, "obm_rk3_substep: U / Gn table is NULL");
    OBM_REQUIRE(Gm || (!has_zeta && !cache_previous), OBM_ENULL, "obm_rk3_substep: G⁻ table is NULL");
    static thread_local SubstepArgs a;
    int rc = make_dims(grid, &a.d, false);
    if (rc) return rc;
    a.has_zeta = has_zeta ? 1 : 0; a.cache = cache_previous ? 1 : 0;
    a.dt = dt; a.gamma = gamma; a.zeta = zeta;
    for (int f0 = 0; f0 < nfields; f0 += TS_MAX_FIELDS) {  // more than 40 tracers: several launches
        a.nfields = nfields - f0 < TS_MAX_FIELDS ? nfields - f0 : TS_MAX_FIELDS;
        for (int f = 0; f < a.nfields; f++) {
            OBM_REQUIRE(U[f0 + f] && Gn[f0 + f] && (!Gm || Gm[f0 + f] || (!has_zeta && !cache_previous)), OBM_ENULL,
                        "obm_rk3_substep: field %d has a NULL pointer", f0 + f);
            a.U[f] = U[f0 + f]; a.Gn[f] = Gn[f0 + f]; a.Gm[f] = Gm ? Gm[f0 + f] : nullptr;
        }
        rk3_substep_kernel<<<cell_grid(a.d, 256), 256, 0, (cudaStream_t)stream>>>(a);
        rc = launch_status("rk3_substep_kernel");
        if (rc) return rc;
    }
    return 0;
}
