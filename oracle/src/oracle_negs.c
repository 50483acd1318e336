/*
 * oracle_negs.c — ORACLE (test infrastructure, not product): CPU restatement of
 * ScaleNegativeTracers / ZeroNegativeTracers and of the tracer-inventory sums.
 *
 * Follows:
 *   src/Utils/negative_tracers.jl:250-276  scale_for_negs_cpu! (clearest statement; the GPU
 *                                          variant :191-240 is arithmetically identical)
 *   src/Utils/negative_tracers.jl:26-32    ZeroNegativeTracers
 *   src/OceanBioME.jl:169                  a tuple of modifiers is applied one after the other
 *                                          (one launch per conserved group)
 *
 * Pinned by the exact outcomes of test/test_utils.jl:7-40 and test/test_PISCES.jl:129-168
 * (tests/test_oracle_negs.py).
 */
#include "oracle_common.h"

/* one group = one launch of scale_for_negs_cpu! */
int orc_scale_negative_group(const obm_grid* g, int n, double* const* fields, const double* scalefactors,
                             double invalid_fill_value) {
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < g->Nz; k++)
        for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
                int64_t idx = cell_index(g, i, j, k);
                /* `if !immersed_cell(i, j, k, grid)` :253 — grid-fitted bottom: below the bottom-most active cell */
                if (g->bottom_indices_xy && (int64_t)k + 1 < g->bottom_indices_xy[plane_index(g, i, j)]) continue;
                double t = 0.0, p = 0.0;
                for (int f = 0; f < n; f++) {
                    double value = fields[f][idx];
                    double sf = scalefactors[f];
                    t += value * sf;
                    if (value > 0) p += value * sf;
                }
                t = t < 0 ? invalid_fill_value : t;
                for (int f = 0; f < n; f++) {
                    double value = fields[f][idx];
                    double new_value = (!isfinite(value) | (value > 0)) ? value * t / p : 0;
                    fields[f][idx] = new_value;
                }
            }
    return 0;
}

/* all groups of a model, sequentially, in the given order */
int orc_scale_negative_tracers(const obm_grid* g, int ntracers, double* const* tracers, int ngroups,
                               const obm_scale_group* groups, double invalid_fill_value) {
    (void)ntracers;
    for (int q = 0; q < ngroups; q++) {
        double* f[OBM_MAX_GROUP_SIZE];
        for (int m = 0; m < groups[q].n; m++) f[m] = tracers[groups[q].index[m]];
        int rc = orc_scale_negative_group(g, groups[q].n, f, groups[q].scalefactor, invalid_fill_value);
        if (rc) return rc;
    }
    return 0;
}

/* negative_tracers.jl:26-32: parent(tracer) .= max.(0.0, parent(tracer)) — Julia max propagates NaN */
int orc_zero_negative_tracers(int64_t n_parent, int ntracers, double* const* tracers) {
    for (int t = 0; t < ntracers; t++)
        for (int64_t q = 0; q < n_parent; q++) tracers[t][q] = jl_max(0.0, tracers[t][q]);
    return 0;
}

/* inventory: out[g] = Σ_cells Σ_f sf·c_f·V  (serial, k-j-i order, long double accumulator so the
 * oracle is the tighter of the two sums being compared) */
int orc_inventory(const obm_grid* g, int ntracers, const double* const* tracers, int ngroups,
                  const obm_scale_group* groups, const double* cell_volume, double uniform_volume, double* out) {
    (void)ntracers;
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
    for (int q = 0; q < ngroups; q++) {
        long double acc = 0.0L;
        for (int k = 0; k < g->Nz; k++)
            for (int j = j0; j < j1; j++)
                for (int i = i0; i < i1; i++) {
                    int64_t idx = cell_index(g, i, j, k);
                    double V = cell_volume ? cell_volume[idx] : uniform_volume;
                    double s = 0.0;
                    for (int m = 0; m < groups[q].n; m++) s += groups[q].scalefactor[m] * tracers[groups[q].index[m]][idx];
                    acc += (long double)(s * V);
                }
        out[q] = (double)acc;
    }
    return 0;
}
