/*
 * oracle_sinking.c — CPU ORACLE (test infrastructure only, see oracle_common.h) for the vertical advection of the
 * sinking tracers by their drift velocity,  G[c] += −∂z(w c)  in flux form.
 *
 * PARITY UNPINNED against Oceananigans: the operator (`div_Uc` → `_advective_tracer_flux_z` with the model's
 * advection scheme, applied to `biogeochemical_drift_velocity(bgc, Val(c)).w`) lives in Oceananigans, which is not
 * in the reference tree (OceanBioME only builds the w fields: src/Utils/sinking_velocity_fields.jl:10-35, and reads
 * the same bottom-face flux in src/Sediments/compute_tendencies.jl).  Restated from the published schemes:
 *   face k lies between cells k−1 and k;  F_k = w_k · c̃_k;  G_k −= (F_{k+1} − F_k) / Δz_k
 *   UpwindBiased(order=1): c̃ = c[k−1] if w > 0 else c[k]
 *   Centered(order=2):     c̃ = (c[k−1] + c[k]) / 2
 *   UpwindBiased(order=3): c̃ = (−c[k−2] + 5c[k−1] + 2c[k]) / 6 if w > 0 else (2c[k−1] + 5c[k] − c[k+1]) / 6,
 *                          first order where that stencil would leave the interior.
 * Pinned instead by properties: telescoping (column integral changes by the boundary fluxes only), exactness of the
 * face values for polynomials of the scheme's degree, and the total-nitrogen conservation of the reference's
 * sediment test (test/test_sediments.jl:37-80) when the bottom face feeds the sediment.
 */
#include "oracle_common.h"

static double face_value(int scheme, const double* col, int64_t sz, int k, int Nz, double w) {
    double below = col[sz * (k - 1)], above = col[sz * k];
    if (scheme == OBM_ADV_CENTERED2) return (below + above) / 2;
    if (scheme == OBM_ADV_UPWIND3) {
        if (w > 0 && k - 2 >= 0 && k <= Nz - 1) return (-col[sz * (k - 2)] + 5 * below + 2 * above) / 6;
        if (w < 0 && k - 1 >= 0 && k + 1 <= Nz - 1) return (2 * below + 5 * above - col[sz * (k + 1)]) / 6;
    }
    return w > 0 ? below : above;
}

int orc_sinking_tendencies(const obm_grid* g, int ntracers, const double* const* tracers, const double* const* w_faces,
                           double* const* G, int advection, int accumulate) {
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
    int64_t sy = (int64_t)g->Nx + 2 * g->Hx, sz = sy * ((int64_t)g->Ny + 2 * g->Hy);
    const double* zf = g->zf + g->Hz;
    for (int t = 0; t < ntracers; t++)
        for (int k = 0; k < g->Nz; k++)
            for (int j = j0; j < j1; j++)
                for (int i = i0; i < i1; i++) {
                    int64_t idx = cell_index(g, i, j, k);
                    const double* col = tracers[t] + (idx - sz * k);
                    double w_lo = w_faces[t][idx], w_hi = w_faces[t][idx + sz];
                    double F_lo = w_lo == 0.0 ? 0.0 : w_lo * face_value(advection, col, sz, k, g->Nz, w_lo);
                    double F_hi = w_hi == 0.0 ? 0.0 : w_hi * face_value(advection, col, sz, k + 1, g->Nz, w_hi);
                    double gg = -(F_hi - F_lo) * (1.0 / (zf[k + 1] - zf[k]));
                    G[t][idx] = accumulate ? G[t][idx] + gg : gg;
                }
    return 0;
}
