/*
 * oracle_timestep.c — CPU ORACLE (test infrastructure only, see oracle_common.h) for the tracer update of
 * src/BoxModel/timesteppers.jl:20-28 (cache_previous_tendencies!) and :66-93 (rk3_substep!):
 *     U += Δt * (γⁿ * Gⁿ + ζⁿ * G⁻)        (first stage, ζ = nothing:  U += Δt * γ¹ * G¹)
 * one pass per tracer like the reference; pure arithmetic identity, nothing to pin beyond operation order.
 */
#include "oracle_common.h"

int orc_rk3_substep(const obm_grid* g, int nfields, double* const* U, const double* const* Gn, double* const* Gm,
                    double dt, double gamma, double zeta, int has_zeta, int cache_previous) {
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
    for (int f = 0; f < nfields; f++)
        for (int k = 0; k < g->Nz; k++)
            for (int j = j0; j < j1; j++)
                for (int i = i0; i < i1; i++) {
                    int64_t idx = cell_index(g, i, j, k);
                    double gn = Gn[f][idx];
                    if (has_zeta) U[f][idx] += dt * (gamma * gn + zeta * Gm[f][idx]);
                    else U[f][idx] += dt * gamma * gn;
                    if (cache_previous) Gm[f][idx] = gn;
                }
    return 0;
}
