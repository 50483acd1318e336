/*
 * oracle_gas_exchange.c — CPU ORACLE (test infrastructure only, see oracle_common.h) for the
 * air–sea gas-exchange flux of OceanBioME.jl v0.17.6, src/Models/GasExchange/.
 *
 * Pinned by the reference's own known answers (test/test_gasexchange_carbon_chem.jl):
 *   :31      DIC flux ≈ −8e-6 ± 1e-6 at T=15, S=35, DIC=2220, Alk=2500, air 413, u₁₀ = 2
 *   :164-165 Sc_CO₂(20 °C) ≈ 668 ± 1, Sc_O₂(20 °C) ≈ 568 ± 1
 *   :180     pCO₂ ≈ 350 ± 0.1 at T=25, S=35, DIC=2136.242890518708, Alk=2500
 *   :184     O₂ air-side value ≈ 200 ± 50
 * (tests/test_oracle_gas_exchange.py).
 */
#include "oracle_common.h"

double orc_carbon_chemistry(double DIC, double T, double S, double Alk, int has_pH, double pH, int has_P, double P,
                            double silicate, double phosphate, double initial_pH_guess, int output_kind, int* n_iters,
                            int* n_fevals);
double orc_K0(double T, double S);
double orc_teos10_polynomial_approximation(double T, double Sp, double Pbar);

/* Julia x^4 for Float64 and a literal exponent: Base.Math.pow_body (compensated power by
 * squaring); x^2 and x^3 lower to x*x and x*x*x (Base.literal_pow). */
static double jl_pow4(double x) {
    double x2 = x * x, l2 = fma(x, x, -x2);
    double err = x2 * 2 * l2;
    double x4 = x2 * x2, l4 = fma(x2, x2, -x4);
    l4 += err;
    return (isfinite(x4) && isfinite(l4)) ? x4 + l4 : x4;
}

/* PolynomialParameterisation{N} — generic_parameterisations.jl:24-36 (the "fast" forms N ≤ 5,
 * n-ary + evaluated left to right) */
double orc_polynomial(int order, const double* c, double x) {
    double y = c[0];
    if (order >= 1) y = y + c[1] * x;
    if (order >= 2) y = y + c[2] * (x * x);
    if (order >= 3) y = y + c[3] * (x * x * x);
    if (order >= 4) y = y + c[4] * jl_pow4(x);
    return y;
}

/* Wanninkhof92Solubility surface value — gas_solubility.jl:34-47.  As in the reference the
 * quadratic term reuses B2 (B3 is never read). */
double orc_w92_solubility(const double* w, double T, double S) {
    double Tk = T + 273.15;
    double Tk_100 = Tk / 100.0;
    double beta = exp(w[0] + w[1] / Tk_100 + w[2] * log(Tk_100) + S * (w[3] + w[4] * Tk_100 + w[4] * (Tk_100 * Tk_100)));
    return beta / Tk;
}

/* SchmidtScaledTransferVelocity — gas_transfer_velocity.jl:32-33 */
double orc_transfer_velocity(const obm_gas_exchange_params* p, double u10, double T, double S) {
    double sol = 1.0;
    if (p->solubility_kind == OBM_GE_SOLUBILITY_K0_RHO) /* gas_solubility.jl:65, density at Pbar = 0 */
        sol = orc_K0(T + 273.15, S) * orc_teos10_polynomial_approximation(T, S, 0.0) / 1000.0;
    return orc_polynomial(p->k660_order, p->k660, u10) / sqrt(orc_polynomial(4, p->schmidt, T) / 660.0) * sol;
}

/* (g::GasExchange)(i, j, grid, clock, model_fields) — gas_exchange.jl:26-38 — one column */
double orc_gas_exchange_point(const obm_gas_exchange_params* p, double T, double S, double tracer, double DIC,
                              double Alk, double silicate, double phosphate, double u10, double air) {
    double k = orc_transfer_velocity(p, u10, T, S);
    if (p->air_kind == OBM_GE_AIR_WANNINKHOF92) air = air * orc_w92_solubility(p->w92, T, S); /* gas_solubility.jl:24-25 */
    double water = tracer;
    if (p->water_kind == OBM_GE_WATER_PCO2) { /* carbon_dioxide_concentration.jl:48-60 */
        double pH0 = p->carbon_chemistry.initial_pH_guess > 0 ? p->carbon_chemistry.initial_pH_guess : 8.0;
        water = orc_carbon_chemistry(DIC, T, S, Alk, 0, 0.0, 0, 0.0, silicate, phosphate, pH0, OBM_CC_PCO2, NULL, NULL);
    }
    return k * (water - air);
}

int orc_gas_exchange_flux(const obm_grid* g, const obm_gas_exchange_params* p, const double* T, const double* S,
                          const double* tracer, const double* DIC, const double* Alk, const double* silicate_f,
                          const double* phosphate_f, const double* wind_xy, const double* air_xy, double* flux_xy,
                          double* G_top) {
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
    const int kt = g->Nz - 1;
    const double dz = g->zf[kt + 1 + g->Hz] - g->zf[kt + g->Hz];
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = j0; j < j1; j++)
        for (int i = i0; i < i1; i++) {
            int64_t idx = cell_index(g, i, j, kt), pidx = plane_index(g, i, j);
            double sil = 0.0, phos = 0.0;
            if (p->use_silicate_phosphate) {
                sil = silicate_f ? silicate_f[idx] : p->silicate;
                phos = phosphate_f ? phosphate_f[idx] : p->phosphate;
            }
            double f = orc_gas_exchange_point(p, T[idx], S[idx], tracer ? tracer[idx] : 0.0, DIC ? DIC[idx] : 0.0,
                                              Alk ? Alk[idx] : 0.0, sil, phos, wind_xy ? wind_xy[pidx] : p->wind_speed,
                                              air_xy ? air_xy[pidx] : p->air_concentration);
            if (flux_xy) flux_xy[pidx] = f;
            if (G_top) G_top[idx] -= f / dz;
        }
    return 0;
}
