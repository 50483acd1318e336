/*
 * oracle_light.c — ORACLE (test infrastructure, not product): CPU restatement of the light
 * attenuation column kernels of OceanBioME.jl v0.17.6, one serial top-down loop per column and
 * one pass per band, exactly like the reference launches them.
 *
 * Follows:
 *   src/Light/2band.jl:1-33                  two-band (Karleskind) PAR
 *   src/Light/multi_band.jl:136-163          Morel band averaging + N-band PAR kernel
 *   src/Light/morel_coefficients.jl:1-31     coefficient tables
 *   src/Light/compute_euphotic_depth.jl:3-29 euphotic depth
 *   src/Models/AdvectedPopulations/PISCES/mean_mixed_layer_properties.jl:25-49  mixed-layer mean
 *
 * Pinned by the reference's analytic known answers: test/test_light.jl:10-106 and
 * test/test_PISCES.jl:98-127 (tests/test_oracle_light.py).
 */
#include "oracle_common.h"

/* 2band.jl:1-33 */
int orc_par_twoband(const obm_grid* g, const obm_twoband_params* m, const double* P, const double* surface_PAR_xy,
                    double surface_PAR_const, double* PAR) {
    const double kr = m->water_red_attenuation, kb = m->water_blue_attenuation;
    const double xr = m->chlorophyll_red_attenuation, xb = m->chlorophyll_blue_attenuation;
    const double er = m->chlorophyll_red_exponent, eb = m->chlorophyll_blue_exponent;
    const double r = m->pigment_ratio, Rcp = m->phytoplankton_chlorophyll_ratio;
    const double* zc = g->zc + g->Hz;
    const double* zf = g->zf + g->Hz;
    const int Nz = g->Nz;
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int j = j0; j < j1; j++)
        for (int i = i0; i < i1; i++) {
            double PAR0 = surface_PAR_xy ? surface_PAR_xy[plane_index(g, i, j)] : surface_PAR_const;
            /* first point below surface (Julia k = Nz ↔ C k = Nz-1) */
            int k = Nz - 1;
            double Pk = P[cell_index(g, i, j, k)];
            double ichl_r = (zf[Nz] - zc[k]) * pow(Pk * Rcp / r, er);
            double ichl_b = (zf[Nz] - zc[k]) * pow(Pk * Rcp / r, eb);
            PAR[cell_index(g, i, j, k)] = PAR0 * (exp(kr * zc[k] - xr * ichl_r) + exp(kb * zc[k] - xb * ichl_b)) / 2;
            for (k = Nz - 2; k >= 0; k--) {
                double Pk1 = P[cell_index(g, i, j, k + 1)];
                Pk = P[cell_index(g, i, j, k)];
                ichl_r += (zc[k + 1] - zf[k + 1]) * pow(Pk1 * Rcp / r, er) + (zf[k + 1] - zc[k]) * pow(Pk * Rcp / r, er);
                ichl_b += (zc[k + 1] - zf[k + 1]) * pow(Pk1 * Rcp / r, eb) + (zf[k + 1] - zc[k]) * pow(Pk * Rcp / r, eb);
                PAR[cell_index(g, i, j, k)] = PAR0 * (exp(kr * zc[k] - xr * ichl_r) + exp(kb * zc[k] - xb * ichl_b)) / 2;
            }
        }
    return 0;
}

/* multi_band.jl:147-163 for ONE band (the reference launches once per band, :170-180).
 * Chl = chl_scale * (chl_a [+ chl_b]) evaluated lazily per access like the reference's
 * AbstractOperation (PISCES/coupling_utils.jl:7; NutrientsPlanktonDetritus/coupling_utils.jl:54). */
int orc_par_multiband_band(const obm_grid* g, double kw, double e, double chi, double surface_PAR_division,
                           const double* chl_a, const double* chl_b, double chl_scale, const double* surface_PAR_xy,
                           double surface_PAR_const, double* field) {
    const double* zc = g->zc + g->Hz;
    const int Nz = g->Nz;
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int j = j0; j < j1; j++)
        for (int i = i0; i < i1; i++) {
            double sPAR = surface_PAR_xy ? surface_PAR_xy[plane_index(g, i, j)] : surface_PAR_const;
            int k = Nz - 1;
            int64_t idx = cell_index(g, i, j, k);
            double Chl = chl_b ? chl_scale * (chl_a[idx] + chl_b[idx]) : (chl_scale * chl_a[idx]);
            field[idx] = sPAR * surface_PAR_division * exp(zc[k] * (kw + chi * pow(Chl, e)));
            for (k = Nz - 2; k >= 0; k--) {
                double dz = zc[k] - zc[k + 1];
                idx = cell_index(g, i, j, k);
                Chl = chl_b ? chl_scale * (chl_a[idx] + chl_b[idx]) : (chl_scale * chl_a[idx]);
                field[idx] = field[cell_index(g, i, j, k + 1)] * exp(dz * (kw + chi * pow(Chl, e)));
            }
        }
    return 0;
}

int orc_par_multiband(const obm_grid* g, const obm_multiband_params* m, const double* chl_a, const double* chl_b,
                      double chl_scale, const double* surface_PAR_xy, double surface_PAR_const, double* const* PAR_bands,
                      double* PAR_total) {
    for (int n = 0; n < m->nbands; n++) {
        int rc = orc_par_multiband_band(g, m->water_attenuation_coefficient[n], m->chlorophyll_exponent[n],
                                        m->chlorophyll_attenuation_coefficient[n], m->surface_PAR_division[n], chl_a, chl_b,
                                        chl_scale, surface_PAR_xy, surface_PAR_const, PAR_bands[n]);
        if (rc) return rc;
    }
    if (PAR_total) { /* total_PAR = sum(fields): ((PAR₁ + PAR₂) + PAR₃) … multi_band.jl:120 */
        int i0, i1, j0, j1;
        grid_range(g, &i0, &i1, &j0, &j1);
        for (int k = 0; k < g->Nz; k++)
            for (int j = j0; j < j1; j++)
                for (int i = i0; i < i1; i++) {
                    int64_t idx = cell_index(g, i, j, k);
                    double s = PAR_bands[0][idx];
                    for (int n = 1; n < m->nbands; n++) s += PAR_bands[n][idx];
                    PAR_total[idx] = s;
                }
    }
    return 0;
}

/* multi_band.jl:97-104,136-140 — band-averaged Morel coefficients.
 * lambda/C are the base tables (n entries); band = [lo, hi] nm. */
double orc_numerical_mean(const double* lambda, const double* C, int n, double lo, double hi) {
    int idx1 = -1, idx2 = -1; /* findlast(base_bands .<= band) */
    for (int q = 0; q < n; q++) {
        if (lambda[q] <= lo) idx1 = q;
        if (lambda[q] <= hi) idx2 = q;
    }
    double integral = 0.0; /* sum([...]) — pairwise in Julia only for n >= 1024 elements; here left to right */
    for (int q = idx1 + 1; q <= idx2; q++) integral += (C[q] + C[q - 1]) * (lambda[q] - lambda[q - 1]) / 2;
    return integral / (lambda[idx2] - lambda[idx1]);
}

/* compute_euphotic_depth.jl:3-29.  Reads PAR[i,j,Nz+1] (halo, :6) as found. */
int orc_euphotic_depth(const obm_grid* g, const double* PAR, double cutoff, double* zeu_xy) {
    const double* zc = g->zc + g->Hz;
    const int Nz = g->Nz;
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
    for (int j = j0; j < j1; j++)
        for (int i = i0; i < i1; i++) {
            double surface_PAR = (PAR[cell_index(g, i, j, Nz - 1)] + PAR[cell_index(g, i, j, Nz)]) / 2;
            double zeu = -INFINITY;
            for (int k = Nz - 2; k >= 0; k--) {
                double PARk = PAR[cell_index(g, i, j, k)];
                if ((PARk <= surface_PAR * cutoff) && isinf(zeu)) {
                    double PARk1 = PAR[cell_index(g, i, j, k + 1)];
                    double zk = zc[k], zk1 = zc[k + 1];
                    zeu = zk + (log(surface_PAR * cutoff) - log(PARk)) * (zk - zk1) / (log(PARk) - log(PARk1));
                }
            }
            zeu_xy[plane_index(g, i, j)] = isfinite(zeu) ? zeu : zc[-1]; /* znode(i,j,0,…) */
        }
    return 0;
}

/* mean_mixed_layer_properties.jl:25-49 */
int orc_mixed_layer_mean(const obm_grid* g, const double* mixed_layer_depth_xy, const double* C, double* Cmxl_xy) {
    const double* zf = g->zf + g->Hz;
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
    for (int j = j0; j < j1; j++)
        for (int i = i0; i < i1; i++) {
            double zmxl = mixed_layer_depth_xy[plane_index(g, i, j)];
            double acc = 0;
            double integration_depth = 0;
            for (int k = g->Nz - 1; k >= 0; k--) {
                double zk = zf[k], zk1 = zf[k + 1];
                double dzk = zk1 - zk;
                double dzk1 = zk1 > zmxl ? zk1 - zmxl : 0;
                double dz = zk >= zmxl ? dzk : dzk1;
                acc += C[cell_index(g, i, j, k)] * dz;
                integration_depth += dz;
            }
            Cmxl_xy[plane_index(g, i, j)] = acc / integration_depth;
        }
    return 0;
}
