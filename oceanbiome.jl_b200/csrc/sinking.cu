// sinking.cu — vertical advection of the sinking tracers by their biogeochemical drift velocity:
//     Gⁿ[c] += −∂z(w c)      for every tracer with `biogeochemical_drift_velocity(bgc, Val(c)).w ≠ 0`,
// all of them in ONE launch.  In the reference this term is part of Oceananigans' `div_Uc(i, j, k, grid, advection,
// total_velocities, c)` inside every tracer's `compute_Gc!` launch; the drift velocities are the z-face fields built
// by `setup_velocity_fields` (src/Utils/sinking_velocity_fields.jl:10-35) / `DepthDependantSinkingSpeed`
// (PISCES/common.jl:39-55).  With no resolved vertical velocity (column / box ensembles — BASELINE config C2 — and
// the reference's sediment tests) this kernel IS the tracer's advection; under a 3-D flow the upwind direction
// depends on w_fluid + w_sink, so Oceananigans keeps the term.
//
// Flux form, face k between cells k−1 and k:  F_k = w_k · c̃_k,   G_k −= (F_{k+1} − F_k) / Δz_k, which telescopes:
// the column integral changes only by the boundary-face fluxes — exactly what the sediment reads at the bottom
// (sediments.cu: same face value) and 0 at the closed top face (w = 0 there).
// Face reconstructions (Oceananigans' uniform-grid coefficients): UpwindBiased(order = 1), Centered(order = 2),
// UpwindBiased(order = 3) — the latter drops to order 1 where its stencil would leave the interior
// (bounded-direction buffer scheme) — and WENO(order = 5) with the Z weights (→ WENO3 → order 1 towards the boundaries).  Tracer halos below / above the column are read as found for orders 1, 2.
// HBM-bound: per tracer per cell 8 B (c; the z-neighbours come from L2) + 8 B (w) + 16 B (Gⁿ read-modify-write).
#include "obm_common.cuh"

namespace obm {

struct SinkArgs {
    GridDims d;
    int n, scheme, accumulate;
    const double* c[OBM_MAX_SINKING_TRACERS];
    const double* w[OBM_MAX_SINKING_TRACERS];
    double* G[OBM_MAX_SINKING_TRACERS];
};

// WENO reconstructions of the face value from the UPWIND side: u0 is the upwind cell next to the face, um1, um2 the two
// behind it, up1, up2 the two ahead (downwind).  Candidate polynomials and smoothness indicators of Jiang & Shu (1996),
// non-linear weights of WENO-Z (Borges et al. 2008): α_k = d_k (1 + (τ / (β_k + ε))²), τ₅ = |β₀ − β₂|, τ₃ = |β₀ − β₁|.
__device__ __forceinline__ double weno5_face(double um2, double um1, double u0, double up1, double up2) {
    const double eps = 1e-8;
    const double p0 = (2 * um2 - 7 * um1 + 11 * u0) / 6, p1 = (-um1 + 5 * u0 + 2 * up1) / 6, p2 = (2 * u0 + 5 * up1 - up2) / 6;
    const double a0 = um2 - 2 * um1 + u0, b0 = um2 - 4 * um1 + 3 * u0;
    const double a1 = um1 - 2 * u0 + up1, b1 = um1 - up1;
    const double a2 = u0 - 2 * up1 + up2, b2 = 3 * u0 - 4 * up1 + up2;
    const double be0 = 13.0 / 12 * (a0 * a0) + 0.25 * (b0 * b0);
    const double be1 = 13.0 / 12 * (a1 * a1) + 0.25 * (b1 * b1);
    const double be2 = 13.0 / 12 * (a2 * a2) + 0.25 * (b2 * b2);
    const double tau = fabs(be0 - be2);
    const double r0 = tau / (be0 + eps), r1 = tau / (be1 + eps), r2 = tau / (be2 + eps);
    const double w0 = 0.1 * (1 + r0 * r0), w1 = 0.6 * (1 + r1 * r1), w2 = 0.3 * (1 + r2 * r2);
    return (w0 * p0 + w1 * p1 + w2 * p2) / (w0 + w1 + w2);
}
__device__ __forceinline__ double weno3_face(double um1, double u0, double up1) {
    const double eps = 1e-8;
    const double p0 = (-um1 + 3 * u0) / 2, p1 = (u0 + up1) / 2;
    const double be0 = (u0 - um1) * (u0 - um1), be1 = (up1 - u0) * (up1 - u0);
    const double tau = fabs(be0 - be1);
    const double r0 = tau / (be0 + eps), r1 = tau / (be1 + eps);
    const double w0 = (1.0 / 3) * (1 + r0 * r0), w1 = (2.0 / 3) * (1 + r1 * r1);
    return (w0 * p0 + w1 * p1) / (w0 + w1);
}

// value of c at face k (0 … Nz) seen by a flow of vertical velocity w; `col` points at cell k = 0 of the column
__device__ __forceinline__ double face_value(int scheme, const double* col, long long sz, int k, int Nz, double w) {
    const double below = col[sz * (k - 1)], above = col[sz * k];
    if (scheme == OBM_ADV_CENTERED2) return (below + above) / 2;
    if (scheme == OBM_ADV_WENO5) {
        // s = +1: the flow comes from below (upwind cell k − 1), s = −1 from above (upwind cell k); m-th cell behind the
        // upwind one: u − s·m, ahead of it: u + s·m.  Order drops where the stencil would leave the interior.
        const int s = w > 0 ? 1 : -1, u = w > 0 ? k - 1 : k;
        const int lo5 = u - 2 * s < u + 2 * s ? u - 2 * s : u + 2 * s, hi5 = u - 2 * s < u + 2 * s ? u + 2 * s : u - 2 * s;
        if (lo5 >= 0 && hi5 <= Nz - 1)
            return weno5_face(col[sz * (u - 2 * s)], col[sz * (u - s)], col[sz * u], col[sz * (u + s)], col[sz * (u + 2 * s)]);
        if (u - 1 >= 0 && u + 1 <= Nz - 1) return weno3_face(col[sz * (u - s)], col[sz * u], col[sz * (u + s)]);
    }
    if (scheme == OBM_ADV_UPWIND3) {
        if (w > 0 && k - 2 >= 0 && k <= Nz - 1) return (-col[sz * (k - 2)] + 5 * below + 2 * above) / 6;
        if (w < 0 && k - 1 >= 0 && k + 1 <= Nz - 1) return (2 * below + 5 * above - col[sz * (k + 1)]) / 6;
    }
    return w > 0 ? below : above;  // first-order upwind: ((w + |w|) c[k−1] + (w − |w|) c[k]) / 2w
}

__global__ void __launch_bounds__(256) sinking_tendency_kernel(const __grid_constant__ SinkArgs a) {
    int i, j, k;
    if (!thread_cell(a.d, i, j, k)) return;
    const long long idx = cell_index(a.d, i, j, k);
    const long long base = idx - a.d.sz * k;  // cell k = 0 of this column
    const double inv_dz = 1.0 / (a.d.zf[k + 1] - a.d.zf[k]);
    for (int t = 0; t < a.n; t++) {
        const double w_lo = a.w[t][idx], w_hi = a.w[t][idx + a.d.sz];
        const double* col = a.c[t] + base;
        // a face whose velocity is exactly 0 carries no flux whatever the (possibly unfilled) halo holds
        const double F_lo = w_lo == 0.0 ? 0.0 : w_lo * face_value(a.scheme, col, a.d.sz, k, a.d.Nz, w_lo);
        const double F_hi = w_hi == 0.0 ? 0.0 : w_hi * face_value(a.scheme, col, a.d.sz, k + 1, a.d.Nz, w_hi);
        const double g = -(F_hi - F_lo) * inv_dz;
        a.G[t][idx] = a.accumulate ? a.G[t][idx] + g : g;
    }
}

}  // namespace obm

using namespace obm;

extern "C" int obm_sinking_tendencies(const obm_grid* grid, int ntracers, const double* const* tracers,
                                      const double* const* w_faces, double* const* G, int advection, int accumulate,
                                      void* stream) {
    OBM_REQUIRE(ntracers >= 0 && ntracers <= OBM_MAX_SINKING_TRACERS, OBM_ESIZE,
                "obm_sinking_tendencies: ntracers = %d outside [0, %d]", ntracers, OBM_MAX_SINKING_TRACERS);
    if (ntracers == 0) return 0;
    OBM_REQUIRE(tracers && w_faces && G, OBM_ENULL, "obm_sinking_tendencies: tracers / w_faces / G is NULL");
    OBM_REQUIRE(advection == OBM_ADV_UPWIND1 || advection == OBM_ADV_CENTERED2 || advection == OBM_ADV_UPWIND3 || advection == OBM_ADV_WENO5, OBM_EENUM,
                "obm_sinking_tendencies: unknown advection scheme %d", advection);
    SinkArgs a;
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    OBM_REQUIRE(grid->Hz >= 1, OBM_ESIZE, "obm_sinking_tendencies: needs at least one halo cell in z (Hz = %d)", grid->Hz);
    a.n = ntracers;
    a.scheme = advection;
    a.accumulate = accumulate ? 1 : 0;
    for (int t = 0; t < ntracers; t++) {
        OBM_REQUIRE(tracers[t] && w_faces[t] && G[t], OBM_ENULL, "obm_sinking_tendencies: field %d has a NULL pointer", t);
        a.c[t] = tracers[t]; a.w[t] = w_faces[t]; a.G[t] = G[t];
    }
    sinking_tendency_kernel<<<cell_grid(a.d, 256), 256, 0, (cudaStream_t)stream>>>(a);
    return launch_status("sinking_tendency_kernel");
}
