/*
 * oracle_kelp.c — CPU ORACLE (test infrastructure only, see oracle_common.h) for the biologically active particles
 * of src/Particles/ with the SugarKelp individual model (src/Models/Individuals/SugarKelp/).
 *
 * Literal restatement: every tendency is its own function that re-evaluates what it needs, exactly like the reference's
 * per-`Val` methods (equations.jl:1-36, coupling.jl:3-57), and the light-inhibition parameter is found by the
 * reference's NewtonRaphsonSolver (Utils/solvers.jl:6-22) with its own tolerance atol = eps(1e-9) and 1000-iteration cap.
 * The drivers follow update_tracer_tendencies.jl:1-48 (one pass per coupled tracer, NearestPoint deposit divided by the
 * cell volume) and tendencies.jl:3-35 + time_stepping.jl:18-48 (all tendencies from the un-stepped state, then Euler).
 *
 * Pinned by the reference's own test (test/test_sugar_kelp.jl:21-90: kelp + tracer nitrogen and carbon conserved over
 * ten steps at t = 60 days, kelp fields change) and by the conservation identities between equations.jl and
 * coupling.jl; there is no absolute golden value for a kelp tendency in the reference.  `nearest_node` goes through
 * Oceananigans' `_fractional_indices` / `interpolator` (not in the tree): restated as the nearest cell centre with
 * half-way points going to the upper node, then `get_node` (tracer_interpolation.jl:5-7).
 */
#include "oracle_common.h"

#define DAY 86400.0
typedef obm_sugar_kelp_params kelp_t;

static double current_factor(const kelp_t* k, double u, double v, double w) { /* :185-193 */
    double U = sqrt(u * u + v * v + w * w);
    return k->current_1 * (1 - exp(-U / k->current_3)) + k->current_2;
}
static double potential_ammonia_uptake(const kelp_t* k, double NH4, double u, double v, double w) { /* :86-93 */
    return k->maximum_ammonia_uptake * current_factor(k, u, v, w) * NH4 / (k->ammonia_half_saturation + NH4);
}
static double temperature_limit(const kelp_t* k, double T) { /* :216-227 */
    double Tl = k->lower_optimal, Tu = k->upper_optimal;
    return jl_max(0, k->lower_gradient * (T - Tl) + 1) * (T < Tl) + jl_max(0, k->upper_gradient * (T - Tu) + 1) * (T > Tu)
           + 1.0 * (Tl <= T && T <= Tu);
}
static double area_limitation(const kelp_t* k, double A) { /* :197-204 */
    double r = A / k->growth_rate_adjustment;
    return k->growth_adjustment_1 * exp(-(r * r)) + k->growth_adjustment_2;
}
static double jl_mod(double a, double b) { /* Julia mod: sign of the divisor */
    double m = fmod(a, b);
    return (m != 0 && ((m < 0) != (b < 0))) ? m + b : m;
}
static double day_length(double phi, double n) { /* :245-255 */
    n -= 171;
    double M = jl_mod(356.5291 + 0.98560028 * n, 360);
    double C = 1.9148 * sin(M * M_PI / 180) + 0.02 * sin(2 * M * M_PI / 180) + 0.0003 * sin(3 * M * M_PI / 180);
    double lam = jl_mod(M + C + 180 + 102.9372, 360);
    double delta = asin(sin(lam * M_PI / 180) * sin(23.44 * M_PI / 180));
    double omega = (sin(-0.83 * M_PI / 180) * sin(phi * M_PI / 180) * sin(delta)) / (cos(phi * M_PI / 180) * cos(delta));
    return omega / 180;
}
double orc_kelp_seasonal_limitation(const kelp_t* k, double t) { /* :229-243 */
    double n = floor(jl_mod(t, 364 * DAY) / DAY);
    double phi = k->adapted_latitude;
    double lam = (day_length(phi, n) - day_length(phi, n - 1)) / (day_length(phi, 76) - day_length(phi, 75));
    return k->photoperiod_1 * (1 + jl_sign(lam) * pow(fabs(lam), .5)) + k->photoperiod_2;
}
static double growth(const kelp_t* k, double t, double A, double N, double C, double T, double NH4, double u, double v, double w) {
    double f = temperature_limit(k, T) * area_limitation(k, A) * orc_kelp_seasonal_limitation(k, t); /* :195-203 */
    double kA = k->structural_dry_weight_per_area, Ns = k->structural_nitrogen;
    double j = potential_ammonia_uptake(k, NH4, u, v, w);
    double muNH4 = j / kA / (N + Ns);
    double muN = 1 - k->minimum_nitrogen_reserve / N;
    double muC = 1 - k->minimum_carbon_reserve / C;
    return f * jl_min(muC, jl_max(muN, muNH4));
}
static double nitrate_uptake(const kelp_t* k, double N, double NO3, double u, double v, double w) { /* :60-71 */
    double Nmax = k->maximum_nitrogen_reserve, Nmin = k->minimum_nitrogen_reserve;
    return jl_max(0, k->maximum_nitrate_uptake * current_factor(k, u, v, w) * (Nmax - N) / (Nmax - Nmin) * NO3
                         / (k->nitrate_half_saturation + NO3));
}
static double ammonia_uptake(const kelp_t* k, double t, double A, double N, double C, double T, double NH4, double u, double v,
                             double w) { /* :73-84 */
    double j = potential_ammonia_uptake(k, NH4, u, v, w);
    double mu = growth(k, t, A, N, C, T, NH4, u, v, w);
    return jl_min(j, mu * k->structural_dry_weight_per_area * (N + k->structural_nitrogen));
}
static double maximum_photosynthesis(double a, double b) { /* :119 */
    return a / (log(1 + a / b)) * (a / (a + b)) * pow(b / (a + b), b / a);
}
static double d_maximum_photosynthesis(double a, double b) { /* :121 */
    return (a * pow(b / (b + a), b / a) * ((log(a / b + 1) * (b * b) + a * log(a / b + 1) * b) * log(b / (b + a)) + a * a))
           / (pow(log(a / b + 1), 2) * b * pow(b + a, 2));
}
double orc_kelp_light_inhibition(const kelp_t* k, double Pm, int* iterations_out) { /* :110-117, solvers.jl:6-22 */
    double a = k->photosynthetic_efficiency, Is = k->saturation_irradiance;
    double atol = 2.0679515313825692e-25; /* eps(1e-9) */
    double x = 1e-9;
    int N = 0;
    double fx = maximum_photosynthesis(a, x) - Pm / Is;
    while ((fabs(fx) > atol) & (N < 1000)) {
        fx = maximum_photosynthesis(a, x) - Pm / Is;
        x -= fx / d_maximum_photosynthesis(a, x);
        N += 1;
    }
    if (iterations_out) *iterations_out = N;
    return x;
}
static double photosynthesis(const kelp_t* k, double T, double PAR) { /* :95-108 */
    PAR *= DAY / (3.99e-10 * 545e12);
    double Tk = T + 273.15;
    double Ta = k->photosynthesis_arrhenius_temp, Tal = k->photosynthesis_low_arrhenius_temp, Tah = k->photosynthesis_high_arrhenius_temp;
    double Tp = k->photosynthesis_ref_temp_1, Tpl = k->photosynthesis_ref_temp_1, Tph = k->photosynthesis_high_temp;
    double a = k->photosynthetic_efficiency, Is = k->saturation_irradiance;
    double Pm = k->photosynthesis_at_ref_temp_1 * exp(Ta / Tp - Ta / Tk) / (1 + exp(Tal / Tk - Tal / Tpl) + exp(Tah / Tph - Tah / Tk));
    double b = orc_kelp_light_inhibition(k, Pm, NULL);
    double ps = a * Is / log(1 + a / b);
    return ps * (1 - exp(-a * PAR / ps)) * exp(-b * PAR / ps);
}
static double respiration(const kelp_t* k, double t, double A, double N, double C, double T, double NO3, double NH4, double u,
                          double v, double w, double mu) { /* :135-158 */
    double Ta = k->respiration_arrhenius_temp, T1 = k->respiration_ref_temp_1, Tk = T + 273.15;
    double f = exp(Ta / T1 - Ta / Tk);
    double Jm = k->maximum_nitrate_uptake + k->maximum_ammonia_uptake;
    double J = nitrate_uptake(k, N, NO3, u, v, w) + ammonia_uptake(k, t, A, N, C, T, NH4, u, v, w);
    return f * (k->base_basal_respiration_rate + k->base_activity_respiration_rate * (mu / k->maximum_specific_growth_rate + J / Jm));
}
static double specific_carbon_exudate(const kelp_t* k, double C) { return 1 - exp(k->exudation * (k->minimum_carbon_reserve - C)); }
static double nitrogen_exudate(const kelp_t* k, double C, double T, double PAR) { /* :167-176 */
    return photosynthesis(k, T, PAR) * specific_carbon_exudate(k, C) * 14 / 12 / k->exudation_redfield_ratio;
}
static double erosion(const kelp_t* k, double A) { /* :178-183 */
    double e = exp(k->erosion_exponent * A);
    return k->base_erosion_rate * e / (1 + k->base_erosion_rate * (e - 1));
}

/* kelp(Val(name), t, A, N, C, u, v, w, T, NO₃, NH₄, PAR); name: 0 A, 1 N, 2 C (equations.jl:1-36), then the coupled
 * tracers 3 NO₃, 4 NH₄, 5 DIC, 6 O₂, 7 DOC, 8 DON, 9 bPOC, 10 bPON (coupling.jl:3-57) */
double orc_kelp(const kelp_t* k, int name, double t, double A, double N, double C, double u, double v, double w, double T,
                double NO3, double NH4, double PAR) {
    double kA = k->structural_dry_weight_per_area, Ns = k->structural_nitrogen, Cs = k->structural_carbon;
    switch (name) {
        case 0: return A * (growth(k, t, A, N, C, T, NH4, u, v, w) - erosion(k, A)) / DAY;
        case 1: {
            double J = nitrate_uptake(k, N, NO3, u, v, w) + ammonia_uptake(k, t, A, N, C, T, NH4, u, v, w);
            double e = nitrogen_exudate(k, C, T, PAR);
            double mu = growth(k, t, A, N, C, T, NH4, u, v, w);
            return ((J - e) / kA - mu * (N + Ns)) / DAY;
        }
        case 2: {
            double P = photosynthesis(k, T, PAR);
            double mu = growth(k, t, A, N, C, T, NH4, u, v, w);
            double R = respiration(k, t, A, N, C, T, NO3, NH4, u, v, w, mu);
            double e = specific_carbon_exudate(k, C);
            return ((P * (1 - e) - R) / kA - mu * (C + Cs)) / DAY;
        }
        case 3: return -nitrate_uptake(k, N, NO3, u, v, w) * A / (DAY * 14 * 0.001);
        case 4: return -ammonia_uptake(k, t, A, N, C, T, NH4, u, v, w) * A / (DAY * 14 * 0.001);
        case 5: {
            double P = photosynthesis(k, T, PAR);
            double mu = growth(k, t, A, N, C, T, NH4, u, v, w);
            double R = respiration(k, t, A, N, C, T, NO3, NH4, u, v, w, mu);
            return -(P - R) * A / (DAY * 12 * 0.001);
        }
        case 6: return -orc_kelp(k, 5, t, A, N, C, u, v, w, T, NO3, NH4, PAR);
        case 7: return specific_carbon_exudate(k, C) * photosynthesis(k, T, PAR) * A / (DAY * 12 * 0.001);
        case 8: return orc_kelp(k, 7, t, A, N, C, u, v, w, T, NO3, NH4, PAR) / k->exudation_redfield_ratio;
        case 9: return erosion(k, A) * kA * A * (C + Cs) / (DAY * 12 * 0.001);
        case 10: return erosion(k, A) * kA * A * (N + Ns) / (DAY * 14 * 0.001);
        default: return NAN;
    }
}

/* get_node — tracer_interpolation.jl:5-7 (1-based) */
static int64_t get_node(int topo, int64_t i, int64_t N) {
    if (topo == OBM_TOPO_FLAT) return 1;
    if (topo == OBM_TOPO_BOUNDED) return i < 1 ? 1 : (i > N ? N : i);
    return i < 1 ? N : (i > N ? 1 : i);
}
static int nearest_regular(double x, double x0, double dx, int topo, int N) {
    if (topo == OBM_TOPO_FLAT) return 0;
    double fi = (x - x0) / dx, fl = floor(fi);
    int64_t lo = (int64_t)fl + 1;
    return (int)get_node(topo, (fi - fl) < 0.5 ? lo : lo + 1, N) - 1;
}
static int nearest_z(const obm_grid* g, double z, int topo) {
    if (topo == OBM_TOPO_FLAT) return 0;
    const double* zc = g->zc + g->Hz;
    int N = g->Nz, lo = -1;
    for (int k = 0; k < N; k++)
        if (zc[k] <= z) lo = k;
    double frac;
    if (lo < 0) frac = 1.0;
    else if (lo >= N - 1) frac = 0.0;
    else frac = (z - zc[lo]) / (zc[lo + 1] - zc[lo]);
    return (int)get_node(topo, frac < 0.5 ? lo + 1 : lo + 2, N) - 1;
}
static int64_t particle_cell(const obm_grid* g, const obm_particles* q, int64_t n, double* volume) {
    int i = nearest_regular(q->x[n], q->x0, q->dx, q->topology[0], g->Nx);
    int j = nearest_regular(q->y[n], q->y0, q->dy, q->topology[1], g->Ny);
    int k = nearest_z(g, q->z[n], q->topology[2]);
    const double* zf = g->zf + g->Hz;
    *volume = q->dx * q->dy * (zf[k + 1] - zf[k]);
    return cell_index(g, i, j, k);
}
int64_t orc_particle_cell(const obm_grid* g, const obm_particles* q, int64_t n) {
    double v;
    return particle_cell(g, q, n, &v);
}

/* update_tendencies!(bgc, particles, model) — update_tracer_tendencies.jl:1-48 */
int orc_kelp_update_tendencies(const obm_grid* g, const kelp_t* k, const obm_particles* q, const obm_kelp_tracers* f,
                               double* const* G, double t) {
    for (int c = 0; c < OBM_KELP_NCOUPLED; c++) { /* one launch per coupled tracer */
        if (!G[c]) continue;
        for (int64_t n = 0; n < q->n; n++) {
            double volume;
            int64_t idx = particle_cell(g, q, n, &volume);
            double u = f->u ? f->u[idx] : 0, v = f->v ? f->v[idx] : 0, w = f->w ? f->w[idx] : 0;
            double pt = orc_kelp(k, 3 + c, t, q->A[n], q->N[n], q->C[n], u, v, w, f->T[idx], f->NO3[idx], f->NH4[idx], f->PAR[idx]);
            double sf = q->scalefactors ? q->scalefactors[n] : 1.0;
            double total = sf * pt;
            G[c][idx] += total / volume;
        }
    }
    return 0;
}

/* time_step_particle_fields!(::ForwardEuler, …) — tendencies.jl:3-35, time_stepping.jl:18-48 */
int orc_kelp_step(const obm_grid* g, const kelp_t* k, const obm_particles* q, const obm_kelp_tracers* f, double t, double dt,
                  double* const* tendencies_out) {
    for (int64_t n = 0; n < q->n; n++) {
        double volume, d[3];
        int64_t idx = particle_cell(g, q, n, &volume);
        double u = f->u ? f->u[idx] : 0, v = f->v ? f->v[idx] : 0, w = f->w ? f->w[idx] : 0;
        for (int name = 0; name < 3; name++)
            d[name] = orc_kelp(k, name, t, q->A[n], q->N[n], q->C[n], u, v, w, f->T[idx], f->NO3[idx], f->NH4[idx], f->PAR[idx]);
        if (tendencies_out) { tendencies_out[0][n] = d[0]; tendencies_out[1][n] = d[1]; tendencies_out[2][n] = d[2]; }
        q->A[n] += d[0] * dt;
        q->N[n] += d[1] * dt;
        q->C[n] += d[2] * dt;
    }
    return 0;
}
