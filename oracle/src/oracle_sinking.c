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
 *   WENO(order=5):         Jiang & Shu (1996) candidates and smoothness indicators with the WENO-Z weights of Borges et
 *                          al. (2008), ε = 1e-8, uniform-grid coefficients; WENO3, then first order, where the stencil
 *                          would leave the interior (which of the published WENO variants Oceananigans' `WENO()` is at
 *                          the reference's pinned version cannot be checked here: not in the tree).
 * Pinned instead by properties: telescoping (column integral changes by the boundary fluxes only), exactness of the
 * face values for polynomials of the scheme's degree, and the total-nitrogen conservation of the reference's
 * sediment test (test/test_sediments.jl:37-80) when the bottom face feeds the sediment.
 */
#include "oracle_common.h"

/* WENO reconstructions of the face value from the UPWIND side: u0 is the upwind cell next to the face, um1, um2 the two
 * behind it, up1, up2 the two ahead (downwind).  Candidate polynomials and smoothness indicators of Jiang & Shu (1996),
 * non-linear weights of WENO-Z (Borges et al. 2008): α_k = d_k (1 + (τ / (β_k + ε))²), τ₅ = |β₀ − β₂|, τ₃ = |β₀ − β₁|. */
static double weno5_face(double um2, double um1, double u0, double up1, double up2) {
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
static double weno3_face(double um1, double u0, double up1) {
    const double eps = 1e-8;
    const double p0 = (-um1 + 3 * u0) / 2, p1 = (u0 + up1) / 2;
    const double be0 = (u0 - um1) * (u0 - um1), be1 = (up1 - u0) * (up1 - u0);
    const double tau = fabs(be0 - be1);
    const double r0 = tau / (be0 + eps), r1 = tau / (be1 + eps);
    const double w0 = (1.0 / 3) * (1 + r0 * r0), w1 = (2.0 / 3) * (1 + r1 * r1);
    return (w0 * p0 + w1 * p1) / (w0 + w1);
}

static double face_value(int scheme, const double* col, int64_t sz, int k, int Nz, double w) {
    double below = col[sz * (k - 1)], above = col[sz * k];
    if (scheme == OBM_ADV_CENTERED2) return (below + above) / 2;
    if (scheme == OBM_ADV_WENO5) {
        /* s = +1: the flow comes from below (upwind cell k − 1), s = −1 from above (upwind cell k); m-th cell behind the
         * upwind one: u − s·m, ahead of it: u + s·m.  Order drops where the stencil would leave the interior. */
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
