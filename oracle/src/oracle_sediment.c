/*
 * oracle_sediment.c — ORACLE (test infrastructure, not product): CPU restatement of the bottom-sediment
 * path of OceanBioME.jl v0.17.6, one :xy pass per field like the reference launches them.
 *
 * Follows:
 *   src/Models/Sediments/instant_remineralisation.jl:103-125   burial efficiency a + b (f/(k+f))²
 *   src/Models/Sediments/simple_multi_G.jl:165-427             Soetaert et al. (2000) level-3 model
 *   src/Sediments/tracked_fields.jl:43-71                      K7 bottom-cell gather, K8 sinking flux
 *   src/Sediments/compute_tendencies.jl:5-50                   K9
 *   src/Sediments/timesteppers.jl:15-94                        K10 AB2 / RK3 substeps, tendency cache
 *   src/Sediments/tracer_coupling.jl:3-39                      K11 G[i,j,k_bottom] += flux / Δzᶜᶜᶠ
 *   src/Sediments/bottom_indices.jl:7-17                       K12
 *
 * PARITY UNPINNED: the reference's sediment testset is commented out (test/test_sediments.jl:106-163), and two
 * ingredients live in Oceananigans (not in /root/reference): the face reconstruction inside
 * `advective_tracer_flux_z` and the order in which `time_step!` steps, caches and recomputes tendencies.
 * Assumed here (stated in DESIGN.md): first-order upwind ((w+|w|)C[k−1] + (w−|w|)C[k])/2 or centred second
 * order; per call: step pools with the STORED Gⁿ/G⁻, cache G⁻ ← Gⁿ, recompute Gⁿ from the new state.
 * What survives of the reference's intent — total nitrogen conservation (test_sediments.jl:37-80) — is tested.
 */
#include "oracle_common.h"

#define DAY 86400.0

typedef const obm_sediment_params* SP;

/* ---- InstantRemineralisation: instant_remineralisation.jl:103-125 ---- */
static double ir_burial_efficiency(SP s, double flux) {
    double q = flux / (s->burial_efficiency_half_saturation + flux);
    return s->burial_efficiency_constant1 + s->burial_efficiency_constant2 * (q * q);
}
static double ir_storage(SP s, double flux) { return ir_burial_efficiency(s, flux) * flux; }
static double ir_remineralisation(SP s, double flux) { return (1 - ir_burial_efficiency(s, flux)) * flux; }

/* ---- SimpleMultiG: simple_multi_G.jl:370-427 ---- */
static double reactivity(SP s, double Cs, double Cf) {
    double Cr = s->slow_decay_rate * Cs + s->fast_decay_rate * Cf;
    return Cr / (Cs + Cf + EPS0);
}
static double ammonia_oxidation_fraction(SP s, double Nr, double Cr, double k, double NH4, double O2) {
    const double* q = s->nitrate_oxidation_params;
    double kO2 = s->anoxia_half_saturation;
    double lC = log(Cr * DAY);
    double ln_pNr = (q[0] + q[1] * lC * log(O2) + q[2] * (lC * lC) + q[3] * log(k * DAY) * log(NH4) + q[4] * lC + q[5] * lC * log(NH4));
    double p = exp(ln_pNr) / (Nr * DAY) * O2 / (kO2 + O2);
    return isfinite(p) ? p : 0.0;
}
static double denitrification_fraction(SP s, double Nr, double Cr, double k, double NO3, double O2) {
    (void)Nr;
    const double* q = s->denitrification_params;
    double kO2 = s->anoxia_half_saturation;
    double lC = log(Cr * DAY), lN = log(NO3), lk = log(k * DAY);
    double ln_pCr = (q[0] + q[1] * lC + q[2] * (lN * lN) + q[3] * (lC * lC) + q[4] * (lk * lk) + q[5] * log(O2) * log(k));
    double p = exp(ln_pCr) / (Cr * DAY) * O2 / (kO2 + O2);
    return isfinite(p) ? p : 0.0;
}
static double anoxic_remineralisation_fraction(SP s, double Nr, double Cr, double k, double NO3, double O2) {
    (void)Nr;
    const double* q = s->anoxic_params;
    double lC = log(Cr * DAY), lN = log(NO3);
    double ln_pCr = (q[0] + q[1] * lC + q[2] * (lC * lC) + q[3] * log(k * DAY) + q[4] * log(O2) * log(k) + q[5] * (lN * lN));
    double p = exp(ln_pCr) / (Cr * DAY);
    return isfinite(p) ? p : 0.0;
}
static double solid_deposition_fraction(SP s) { return 0.223 * pow(s->sedimentation_rate, 0.336); }

/* the values one continuous-form sediment callable sees: pools…, tracked tracers…, fluxes… */
typedef struct {
    double pool[OBM_SED_MAX_POOLS];
    double NO3, NH4, O2;
    double fN, fC; /* Σ sinking nitrogen / carbon fluxes (sum(fluxs), left to right) */
} sed_point;

/* remineralisation rates (Nr, Cr) and reactivity k — simple_multi_G.jl:196-203 / :311-317 */
static void smg_rates(SP s, const sed_point* c, double* Nr, double* Cr, double* k) {
    double Ns = c->pool[0], Nf = c->pool[1];
    *Nr = s->slow_decay_rate * Ns + s->fast_decay_rate * Nf;
    if (s->carbon) {
        double Cs = c->pool[3], Cf = c->pool[4];
        *Cr = s->slow_decay_rate * Cs + s->fast_decay_rate * Cf;
        *k = reactivity(s, Cs, Cf);
    } else {
        double R = s->sinking_redfield;
        *Cr = *Nr * R;
        *k = reactivity(s, Ns * R, Nf * R);
    }
}

/* sediment pool tendency `biogeochemistry(Val(name), x, y, t, fields...)` — pool index in required order */
static double pool_tendency(SP s, const sed_point* c, int n) {
    if (s->model == OBM_SED_INSTANT_REMINERALISATION) return ir_storage(s, c->fN);
    double fr = s->refactory_fraction;
    switch (n) {
        case 0: return (1 - fr) * s->slow_fraction * c->fN - s->slow_decay_rate * c->pool[0]; /* Ns :165-173 */
        case 1: return (1 - fr) * s->fast_fraction * c->fN - s->fast_decay_rate * c->pool[1]; /* Nf :175-183 */
        case 2: return fr * c->fN;                                                             /* Nr :185-191 */
        case 3: return (1 - fr) * s->slow_fraction * c->fC - s->slow_decay_rate * c->pool[3]; /* Cs :279-287 */
        case 4: return (1 - fr) * s->fast_fraction * c->fC - s->fast_decay_rate * c->pool[4]; /* Cf :289-297 */
        default: return fr * c->fC;                                                            /* Cr :299-305 */
    }
}

/* coupled tracer flux: InstantRemineralisation: 0 = receiver; SimpleMultiG: 0 NO₃, 1 NH₄, 2 O₂, 3 DIC */
static double coupled_flux(SP s, const sed_point* c, int n) {
    if (s->model == OBM_SED_INSTANT_REMINERALISATION) return ir_remineralisation(s, c->fN);
    double Nr, Cr, k;
    smg_rates(s, c, &Nr, &Cr, &k);
    if (n == 3) return Cr; /* DIC :361-368 */
    double pn = ammonia_oxidation_fraction(s, Nr, Cr, k, c->NH4, c->O2);
    double pnp = denitrification_fraction(s, Nr, Cr, k, c->NO3, c->O2);
    if (n == 1) return (1 - pn) * Nr + 0.8 * pnp * Cr; /* NH₄ :195-209 */
    if (n == 0) return pn * Nr - 0.8 * pnp * Cr;       /* NO₃ :211-226 */
    double pa = anoxic_remineralisation_fraction(s, Nr, Cr, k, c->NO3, c->O2);
    double ps = solid_deposition_fraction(s);
    double kO2 = s->anoxia_half_saturation;
    return -(1 - pa * ps - pnp) * c->O2 / (kO2 + c->O2) * Cr - 2 * pn * Nr; /* O₂ :228-247 */
}

static int npools(SP s) { return s->model == OBM_SED_INSTANT_REMINERALISATION ? 1 : (s->carbon ? 6 : 3); }
static int ncoupled(SP s) { return s->model == OBM_SED_INSTANT_REMINERALISATION ? 1 : (s->carbon ? 4 : 3); }
static int ntracked_tracers(SP s) { return s->model == OBM_SED_INSTANT_REMINERALISATION ? 0 : 3; }

static int64_t kbottom(const obm_grid* g, const obm_sediment_fields* f, int i, int j) {
    return f->bottom_indices_xy ? f->bottom_indices_xy[plane_index(g, i, j)] - 1 : 0; /* 0-based */
}

/* sinking_flux tracked_fields.jl:60-61 with Oceananigans' upwind_biased_product / centred face value */
static double sinking_flux(const obm_grid* g, SP s, const double* C, const double* w, int i, int j, int k) {
    int64_t idx = cell_index(g, i, j, k);
    int64_t sz = ((int64_t)g->Nx + 2 * g->Hx) * ((int64_t)g->Ny + 2 * g->Hy);
    double wk = w[idx], CL = C[idx - sz], CR = C[idx];
    if (s->advection == OBM_ADV_UPWIND1) return -(((wk + fabs(wk)) * CL + (wk - fabs(wk)) * CR) / 2);
    return -(wk * ((CL + CR) / 2));
}

static void gather(const obm_grid* g, SP s, const obm_sediment_fields* f, int i, int j, sed_point* c) {
    int64_t pl = plane_index(g, i, j);
    int k = (int)kbottom(g, f, i, j);
    int64_t idx = cell_index(g, i, j, k);
    for (int n = 0; n < npools(s); n++) c->pool[n] = f->pools[n][pl];
    c->NO3 = c->NH4 = c->O2 = 0;
    if (ntracked_tracers(s)) { c->NO3 = f->NO3[idx]; c->NH4 = f->NH4[idx]; c->O2 = f->O2[idx]; }
    c->fN = c->fC = 0;
    for (int n = 0; n < s->nsinking_nitrogen; n++) {
        double fl = sinking_flux(g, s, f->sinking[n], f->sinking_w[n], i, j, k);
        c->fN = n == 0 ? fl : c->fN + fl;
    }
    for (int n = 0; n < s->nsinking_carbon; n++) {
        int q = s->nsinking_nitrogen + n;
        double fl = sinking_flux(g, s, f->sinking[q], f->sinking_w[q], i, j, k);
        c->fC = n == 0 ? fl : c->fC + fl;
    }
}

/* update_biogeochemical_state!(model, sediment) — Sediments/update_state.jl:6-16 */
int orc_sediment_update_state(const obm_grid* g, const obm_sediment_params* s, const obm_sediment_fields* f, double dt,
                              double chi) {
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
    const int np = npools(s), nt = ntracked_tracers(s);
    for (int j = j0; j < j1; j++)
        for (int i = i0; i < i1; i++) {
            int64_t pl = plane_index(g, i, j);
            sed_point c;
            gather(g, s, f, i, j, &c); /* K7, K8 */
            if (f->tracked_xy[0]) {
                int q = 0;
                if (nt) { f->tracked_xy[q++][pl] = c.NO3; f->tracked_xy[q++][pl] = c.NH4; f->tracked_xy[q++][pl] = c.O2; }
                int k = (int)kbottom(g, f, i, j);
                for (int n = 0; n < s->nsinking_nitrogen + s->nsinking_carbon; n++)
                    f->tracked_xy[q++][pl] = sinking_flux(g, s, f->sinking[n], f->sinking_w[n], i, j, k);
            }
            if (isfinite(dt) && s->timestepper == OBM_TS_RK3) {
                /* time_step!(sediment_model, Δt) with a RungeKutta3TimeStepper: three stages inside the hook, each
                 * rk3_substep! (Sediments/timesteppers.jl:46-73; first stage Δt * γ¹ * G¹) → cache_previous_tendencies!
                 * (:84-96) → update_state! (compute_sediment_tendencies!), tracked fields held fixed.  γ, ζ are
                 * Oceananigans' RungeKutta3TimeStepper constants (not in the tree): 8/15, 5/12, 3/4; −17/60, −5/12. */
                const double gam[3] = {8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0}, zet[3] = {0.0, -17.0 / 60.0, -5.0 / 12.0};
                for (int st = 0; st < 3; st++) {
                    for (int n = 0; n < np; n++) { /* one launch per pool */
                        double Gn = f->Gn[n][pl], Gm = f->Gm[n][pl];
                        if (st == 0) f->pools[n][pl] += dt * gam[0] * Gn;
                        else f->pools[n][pl] += dt * (gam[st] * Gn + zet[st] * Gm);
                    }
                    for (int n = 0; n < np; n++) f->Gm[n][pl] = f->Gn[n][pl];
                    for (int n = 0; n < np; n++) c.pool[n] = f->pools[n][pl];
                    for (int n = 0; n < np; n++) f->Gn[n][pl] = pool_tendency(s, &c, n);
                }
                continue;
            }
            if (isfinite(dt)) { /* time_step!(sediment_model, Δt), QuasiAdamsBashforth2 */
                for (int n = 0; n < np; n++) { /* K10 */
                    double Gn = f->Gn[n][pl], Gm = f->Gm[n][pl], u = f->pools[n][pl];
                    int not_euler = chi != -0.5;
                    double Gu = (1.5 + chi) * Gn - (not_euler ? (0.5 + chi) * Gm : 0.0);
                    u += dt * Gu;
                    f->pools[n][pl] = u;
                    c.pool[n] = u;
                    f->Gm[n][pl] = Gn; /* cache_previous_tendencies! */
                }
            }
            for (int n = 0; n < np; n++) f->Gn[n][pl] = pool_tendency(s, &c, n); /* K9 */
        }
    return 0;
}

/* update_tendencies!(bgc, sediment, model) — Sediments/tracer_coupling.jl:3-39 */
int orc_sediment_update_tendencies(const obm_grid* g, const obm_sediment_params* s, const obm_sediment_fields* f) {
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
    const double* zc = g->zc + g->Hz;
    for (int n = 0; n < ncoupled(s); n++) { /* one launch per coupled tracer */
        if (!f->G_coupled[n]) continue;
        for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
                sed_point c;
                gather(g, s, f, i, j, &c);
                int k = (int)kbottom(g, f, i, j);
                double dz = zc[k] - zc[k - 1]; /* Δzᶜᶜᶠ(i, j, k, grid) */
                f->G_coupled[n][cell_index(g, i, j, k)] += coupled_flux(s, &c, n) / dz;
            }
    }
    return 0;
}

/* K12 find_bottom_cell! — bottom_indices.jl:7-17 with a grid-fitted bottom: immersed ⇔ z_c[k] <= bottom_height */
int orc_find_bottom_cells(const obm_grid* g, const double* bottom_height_xy, int64_t* bottom_indices_xy) {
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
    const double* zc = g->zc + g->Hz;
    for (int j = j0; j < j1; j++)
        for (int i = i0; i < i1; i++) {
            int64_t pl = plane_index(g, i, j);
            int kb = 1;
            while ((zc[kb - 1] <= bottom_height_xy[pl]) && (kb < g->Nz)) kb += 1;
            bottom_indices_xy[pl] = kb;
        }
    return 0;
}

/* scalar entry points for unit tests */
double orc_sediment_pool_tendency(const obm_sediment_params* s, const double* pools, double NO3, double NH4, double O2,
                                  double fN, double fC, int n) {
    sed_point c;
    for (int q = 0; q < OBM_SED_MAX_POOLS; q++) c.pool[q] = pools[q];
    c.NO3 = NO3; c.NH4 = NH4; c.O2 = O2; c.fN = fN; c.fC = fC;
    return pool_tendency(s, &c, n);
}
double orc_sediment_coupled_flux(const obm_sediment_params* s, const double* pools, double NO3, double NH4, double O2,
                                 double fN, double fC, int n) {
    sed_point c;
    for (int q = 0; q < OBM_SED_MAX_POOLS; q++) c.pool[q] = pools[q];
    c.NO3 = NO3; c.NH4 = NH4; c.O2 = O2; c.fN = fN; c.fC = fC;
    return coupled_flux(s, &c, n);
}
