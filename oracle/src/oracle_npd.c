/*
 * oracle_npd.c — ORACLE (test infrastructure, not product): CPU restatement of the
 * Nutrients–Plankton–Detritus family (NPZD, LOBSTER, ± Fe, ± CarbonateSystem, ± Oxygen,
 * 4 detritus choices) of OceanBioME.jl v0.17.6.
 *
 * Follows, function by function and in the reference's operation order,
 *   src/Models/AdvectedPopulations/NutrientsPlanktonDetritus/{nutrients,plankton,detritus,
 *   carbonate_system,oxygen}.jl      (cited per function below as file:line)
 * and keeps the reference's LAUNCH STRUCTURE: one pass over the grid per tracer, each pass
 * re-reading its inputs and re-evaluating every shared intermediate (that is what
 * Oceananigans' per-tracer compute_Gc! does).
 *
 * Parity pinned by: conservation identities of test/test_NutrientsPlanktonDetritus.jl:114-139,
 * zero-state test :101-112 (tests/test_oracle_npd.py).  No absolute tendency goldens exist in
 * the reference — stated in DESIGN.md.
 */
#include <stdio.h>
#include <string.h>

#include "oracle_common.h"

enum {
    R_NO3 = 0, R_NH4, R_FE, R_N, R_P, R_Z, R_T, R_D, R_SPOM, R_BPOM, R_DOM,
    R_SPOC, R_BPOC, R_DOC, R_DIC, R_ALK, R_O2, R_COUNT
};

/* tracer order = required_biogeochemical_tracers — NutrientsPlanktonDetritus.jl:69-74,
 * nutrients.jl:21,46,83; plankton.jl:80; detritus.jl:39,83,293; carbonate_system.jl:39-40;
 * oxygen.jl:19 */
int orc_npd_layout(const obm_npd_params* p, int* roles, char (*names)[16]) {
    int n = 0;
#define ADD(role, nm)                                   \
    do {                                                \
        if (roles) roles[n] = role;                     \
        if (names) {                                    \
            memset(names[n], 0, 16);                    \
            memcpy(names[n], nm, strlen(nm) < 15 ? strlen(nm) : 15); \
        }                                               \
        n++;                                            \
    } while (0)
    switch (p->nutrients) {
        case OBM_NUT_NUTRIENT: ADD(R_N, "N"); break;
        case OBM_NUT_NITRATE_AMMONIA: ADD(R_NO3, "NO₃"); ADD(R_NH4, "NH₄"); break;
        case OBM_NUT_NITRATE_AMMONIA_IRON: ADD(R_NO3, "NO₃"); ADD(R_NH4, "NH₄"); ADD(R_FE, "Fe"); break;
        default: return OBM_EENUM;
    }
    ADD(R_P, "P");
    ADD(R_Z, "Z");
    if (p->has_temperature_coefficient) ADD(R_T, "T");
    switch (p->detritus) {
        case OBM_DET_NONE: break;
        case OBM_DET_DETRITUS: ADD(R_D, "D"); break;
        case OBM_DET_TWO_PARTICLE: ADD(R_SPOM, "sPOM"); ADD(R_BPOM, "bPOM"); ADD(R_DOM, "DOM"); break;
        case OBM_DET_VARIABLE_REDFIELD:
            ADD(R_SPOC, "sPOC"); ADD(R_BPOC, "bPOC"); ADD(R_DOC, "DOC");
            ADD(R_SPOM, "sPON"); ADD(R_BPOM, "bPON"); ADD(R_DOM, "DON");
            break;
        default: return OBM_EENUM;
    }
    int N = p->carbonate_replicates;
    if (N < 0 || 2 * N + n + 1 > OBM_NPD_MAX_TRACERS) return OBM_ESIZE;
    if (N == 1) {
        ADD(R_DIC, "DIC");
        ADD(R_ALK, "Alk");
    } else if (N > 1) {
        char buf[16];
        for (int r = 1; r <= N; r++) { snprintf(buf, 16, "DIC%d", r); ADD(R_DIC, buf); }
        for (int r = 1; r <= N; r++) { snprintf(buf, 16, "Alk%d", r); ADD(R_ALK, buf); }
    }
    if (p->oxygen) ADD(R_O2, "O₂");
#undef ADD
    return n;
}

/* the values one per-point callable can see: fields.X[i,j,k] and auxiliary_fields.PAR[i,j,k] */
typedef struct {
    double v[R_COUNT];
    double PAR;
} cell_t;

typedef const obm_npd_params* P_;
typedef const cell_t* C_;

/* plankton.jl:86-90 */
static double mortality(int form, double X, double m) { return form == OBM_LINEAR ? m * X : m * (X * X); }
static double concentration_limit(int form, double X, double k) {
    return form == OBM_LINEAR ? X / (X + k) : (X * X) / (X * X + k * k);
}

/* detritus.jl:41-42,103-104,295-296,308 */
static double small_particulate_concentration(P_ p, C_ c) {
    switch (p->detritus) {
        case OBM_DET_TWO_PARTICLE:
        case OBM_DET_VARIABLE_REDFIELD: return c->v[R_SPOM];
        case OBM_DET_DETRITUS: return c->v[R_D] * p->small_particle_fraction;
        default: return 0.0;
    }
}
static double large_particulate_concentration(P_ p, C_ c) { (void)p; return c->v[R_BPOM]; }
static double dissolved_organic_nitrogen(P_ p, C_ c) { (void)p; return c->v[R_DOM]; }
/* detritus.jl:50-57,112-119 */
static double small_particulate_carbon_concentration(P_ p, C_ c) {
    return p->detritus == OBM_DET_VARIABLE_REDFIELD ? c->v[R_SPOC] : c->v[R_SPOM] * p->detritus_redfield_ratio;
}
static double large_particulate_carbon_concentration(P_ p, C_ c) {
    return p->detritus == OBM_DET_VARIABLE_REDFIELD ? c->v[R_BPOC] : c->v[R_BPOM] * p->detritus_redfield_ratio;
}
static double dissolved_organic_carbon(P_ p, C_ c) {
    return p->detritus == OBM_DET_VARIABLE_REDFIELD ? c->v[R_DOC] : c->v[R_DOM] * p->detritus_redfield_ratio;
}

/* plankton.jl:393-397 */
static double weighted_phytoplankton_preference(P_ p, double P, double sPOM) {
    double pt = p->preference_for_phytoplankton;
    return pt * P / (pt * P + (1 - pt) * sPOM + EPS0);
}

/* plankton.jl:118-135 */
static double total_grazing(P_ p, C_ c) {
    double kG = p->grazing_half_saturation, g = p->maximum_grazing_rate;
    double Z = c->v[R_Z], P = c->v[R_P];
    double sPOM = small_particulate_concentration(p, c);
    double pr = weighted_phytoplankton_preference(p, P, sPOM);
    double food = pr * P + (1 - pr) * sPOM;
    double L = concentration_limit(p->grazing_concentration_formulation, food, kG);
    return g * L * Z;
}

/* plankton.jl:150-166 */
static double nitrogen_limitation(P_ p, C_ c) {
    double kNO3 = p->nitrate_half_saturation, kNH4 = p->ammonia_half_saturation, psi = p->nitrate_ammonia_inhibition;
    double NO3 = c->v[R_NO3], NH4 = c->v[R_NH4];
    double nitrate_limitation = NO3 * exp(-psi * NH4) / (NO3 + kNO3);
    double ammonia_limitation = jl_max(0.0, NH4 / (kNH4 + NH4));
    return (nitrate_limitation + ammonia_limitation) / 2;
}

/* plankton.jl:168-188 */
static double nutrient_limitation(P_ p, C_ c) {
    switch (p->nutrients) {
        case OBM_NUT_NITRATE_AMMONIA: return nitrogen_limitation(p, c);
        case OBM_NUT_NITRATE_AMMONIA_IRON: {
            double kFe = p->iron_half_saturation, Fe = c->v[R_FE];
            double iron_limitation = Fe / (kFe + Fe);
            return nitrogen_limitation(p, c) * iron_limitation;
        }
        default: {
            double N = c->v[R_N];
            return N / (N + p->nitrate_half_saturation);
        }
    }
}

/* plankton.jl:208-220 */
static double temperature_limitation(P_ p, C_ c) {
    if (!p->has_temperature_coefficient) return 1.0;
    return pow(p->temperature_coefficient, c->v[R_T] / 10);
}
static double light_limitation(P_ p, double PAR, double kPAR) {
    return p->light_limitation == OBM_LIGHT_MONDO ? PAR / (kPAR + PAR) : PAR / sqrt(PAR * PAR + kPAR * kPAR);
}

/* plankton.jl:191-206 */
static double phytoplankton_growth(P_ p, C_ c) {
    double Ln = nutrient_limitation(p, c);
    double Ll = light_limitation(p, c->PAR, p->light_half_saturation);
    double Lt = temperature_limitation(p, c);
    return p->phytoplankton_maximum_growth_rate * Ll * Ln * Lt * c->v[R_P];
}

/* plankton.jl:223-278 */
static double nutrient_uptake_NO3(P_ p, C_ c) {
    double mu = phytoplankton_growth(p, c);
    double NO3 = c->v[R_NO3], NH4 = c->v[R_NH4];
    double nl = NO3 * exp(-p->nitrate_ammonia_inhibition * NH4) / (NO3 + p->nitrate_half_saturation);
    double al = jl_max(0.0, NH4 / (p->ammonia_half_saturation + NH4));
    return mu * nl / (nl + al + EPS0);
}
static double nutrient_uptake_NH4(P_ p, C_ c) {
    double alpha = p->ammonia_fraction_of_exudate, gamma = p->phytoplankton_exudation_fraction;
    double mu = phytoplankton_growth(p, c);
    double NO3 = c->v[R_NO3], NH4 = c->v[R_NH4];
    double nl = NO3 * exp(-p->nitrate_ammonia_inhibition * NH4) / (NO3 + p->nitrate_half_saturation);
    double al = jl_max(0.0, NH4 / (p->ammonia_half_saturation + NH4));
    double waste = alpha * gamma * mu;
    ORC_TERMS(mu * al / (nl + al + EPS0), waste);
    return mu * al / (nl + al + EPS0) - waste;
}
static double nutrient_uptake_Fe(P_ p, C_ c) { return p->iron_ratio * phytoplankton_growth(p, c); }
static double nutrient_uptake_N(P_ p, C_ c) {
    double mu = phytoplankton_growth(p, c);
    return mu * (1 - p->ammonia_fraction_of_exudate * p->phytoplankton_exudation_fraction);
}

/* plankton.jl:280-289 */
static double phytoplankton_primary_production(P_ p, C_ c) {
    double alpha = p->ammonia_fraction_of_exudate, gamma = p->phytoplankton_exudation_fraction;
    double rho = p->carbon_calcite_ratio, R = p->redfield_ratio;
    double muP = phytoplankton_growth(p, c);
    return (1 + rho * (1 - gamma) - alpha * gamma) * muP * R;
}

/* plankton.jl:292-306 */
static double plankton_inorganic_nitrogen_waste(P_ p, C_ c) {
    double aP = p->phytoplankton_solid_waste_fraction, aZ = p->excretion_inorganic_fraction;
    double mu = p->zooplankton_excretion_rate, mP = p->phytoplankton_mortality_rate;
    double P = c->v[R_P], Z = c->v[R_Z];
    double nuP = mortality(p->phytoplankton_mortality_formulation, P, mP);
    return aZ * mu * Z + (1 - aP) * nuP;
}
/* plankton.jl:308-309 */
static double plankton_inorganic_carbon_waste(P_ p, C_ c) { return p->redfield_ratio * plankton_inorganic_nitrogen_waste(p, c); }

/* plankton.jl:312-325 */
static double plankton_organic_nitrogen_waste(P_ p, C_ c) {
    double aZ = p->excretion_inorganic_fraction, mu = p->zooplankton_excretion_rate;
    double aP = p->ammonia_fraction_of_exudate, gamma = p->phytoplankton_exudation_fraction;
    double Z = c->v[R_Z];
    double muP = phytoplankton_growth(p, c);
    return (1 - aP) * gamma * muP + (1 - aZ) * mu * Z;
}
/* plankton.jl:327-328 */
static double plankton_organic_carbon_waste(P_ p, C_ c) { return p->redfield_ratio * plankton_organic_nitrogen_waste(p, c); }

/* plankton.jl:330-346 */
static double solid_waste(P_ p, C_ c) {
    double aP = p->phytoplankton_solid_waste_fraction, aZ = p->zooplankton_assimilation_fraction;
    double mP = p->phytoplankton_mortality_rate, mZ = p->zooplankton_mortality_rate;
    double P = c->v[R_P], Z = c->v[R_Z];
    double G = total_grazing(p, c);
    double nuP = mortality(p->phytoplankton_mortality_formulation, P, mP);
    return (1 - aZ) * G + aP * nuP + mZ * (Z * Z);
}
/* plankton.jl:348-349 */
static double solid_carbon_waste(P_ p, C_ c) { return solid_waste(p, c) * p->redfield_ratio; }

/* plankton.jl:352-388 */
static double grazing_P(P_ p, C_ c) {
    double kG = p->grazing_half_saturation, g = p->maximum_grazing_rate;
    double Z = c->v[R_Z], P = c->v[R_P];
    double sPOM = small_particulate_concentration(p, c);
    double pr = weighted_phytoplankton_preference(p, P, sPOM);
    double food = pr * P + (1 - pr) * sPOM;
    double L = concentration_limit(p->grazing_concentration_formulation, food, kG);
    return g * pr * L * P / (food + jl_eps(food)) * Z;
}
static double grazing_sPOM(P_ p, C_ c) {
    double kG = p->grazing_half_saturation, g = p->maximum_grazing_rate;
    double Z = c->v[R_Z], P = c->v[R_P];
    double sPOM = small_particulate_concentration(p, c);
    double pr = weighted_phytoplankton_preference(p, P, sPOM);
    double food = pr * P + (1 - pr) * sPOM;
    double L = concentration_limit(p->grazing_concentration_formulation, food, kG);
    return g * (1 - pr) * L * sPOM / (food + jl_eps(food)) * Z;
}
/* plankton.jl:390-391 */
static double grazing_sPOC(P_ p, C_ c) { return grazing_sPOM(p, c) * p->redfield_ratio; }

/* plankton.jl:399-412 */
static double calcite_production(P_ p, C_ c) {
    double mP = p->phytoplankton_mortality_rate, R = p->redfield_ratio, rho = p->carbon_calcite_ratio;
    double eta = p->zooplankton_gut_calcite_dissolution;
    double P = c->v[R_P];
    double G = grazing_P(p, c);
    double nu = mortality(p->phytoplankton_mortality_formulation, P, mP);
    return (G * (1 - eta) + nu) * rho * R;
}
/* plankton.jl:414-441 — the fixed-Redfield override applies to TwoParticleAndDissolved,
 * Detritus and Nothing; VariableRedfieldDetritus takes the generic method. */
static double calcite_dissolution(P_ p, C_ c) {
    double R = p->redfield_ratio, rho = p->carbon_calcite_ratio;
    double G = grazing_P(p, c);
    if (p->detritus == OBM_DET_VARIABLE_REDFIELD) {
        double eta = p->zooplankton_gut_calcite_dissolution;
        return G * eta * rho * R;
    }
    double nu = mortality(p->phytoplankton_mortality_formulation, c->v[R_P], p->phytoplankton_mortality_rate);
    return (G + nu) * rho * R;
}
/* plankton.jl:443-450 */
static double calcite_uptake(P_ p, C_ c) {
    double muP = phytoplankton_growth(p, c);
    return 2 * p->carbon_calcite_ratio * muP * p->redfield_ratio;
}

/* detritus.jl:161-207, 298-302, 311-317 */
static double detritus_inorganic_nitrogen_waste(P_ p, C_ c) {
    switch (p->detritus) {
        case OBM_DET_DETRITUS: return c->v[R_D] * p->remineralisation_rate;
        case OBM_DET_NONE: return plankton_organic_nitrogen_waste(p, c) + solid_waste(p, c);
        default: {
            double a = p->remineralisation_inorganic_fraction;
            double sm = p->small_remineralisation_rate, bm = p->large_remineralisation_rate, dm = p->dissolved_remineralisation_rate;
            double sPOM = small_particulate_concentration(p, c), bPOM = large_particulate_concentration(p, c);
            double DOM = dissolved_organic_nitrogen(p, c);
            return (a * (sm * sPOM + bm * bPOM) + dm * DOM);
        }
    }
}
static double detritus_organic_nitrogen_waste(P_ p, C_ c) {
    double a = p->remineralisation_inorganic_fraction;
    double sm = p->small_remineralisation_rate, bm = p->large_remineralisation_rate;
    double sPOM = small_particulate_concentration(p, c), bPOM = large_particulate_concentration(p, c);
    return (1 - a) * (sm * sPOM + bm * bPOM);
}
static double detritus_inorganic_carbon_waste(P_ p, C_ c) {
    switch (p->detritus) {
        case OBM_DET_DETRITUS: return c->v[R_D] * p->remineralisation_rate * p->detritus_redfield_ratio;
        case OBM_DET_NONE: return (plankton_organic_nitrogen_waste(p, c) + solid_waste(p, c)) * p->redfield_ratio;
        default: {
            double a = p->remineralisation_inorganic_fraction;
            double sm = p->small_remineralisation_rate, bm = p->large_remineralisation_rate, dm = p->dissolved_remineralisation_rate;
            double sPOC = small_particulate_carbon_concentration(p, c), bPOC = large_particulate_carbon_concentration(p, c);
            double DOC = dissolved_organic_carbon(p, c);
            return a * (sm * sPOC + bm * bPOC) + dm * DOC;
        }
    }
}
static double detritus_organic_carbon_waste(P_ p, C_ c) {
    double a = p->remineralisation_inorganic_fraction;
    double sm = p->small_remineralisation_rate, bm = p->large_remineralisation_rate;
    double sPOC = small_particulate_carbon_concentration(p, c), bPOC = large_particulate_carbon_concentration(p, c);
    return (1 - a) * (sm * sPOC + bm * bPOC);
}

/* nutrients.jl:29,91 */
static double nitrification(P_ p, C_ c) {
    return p->nutrients == OBM_NUT_NUTRIENT ? 0.0 : p->nitrification_rate * c->v[R_NH4];
}

static int has_NA(P_ p) { return p->nutrients != OBM_NUT_NUTRIENT; }

/* The per-tracer callable `bgc(i,j,k,grid,Val(name),clock,fields,auxiliary_fields)`.
 * Any (model, name) pair without a method falls to `zero(grid)`
 * (NutrientsPlanktonDetritus.jl:88). */
static double tendency(P_ p, C_ c, int role) {
    switch (role) {
        case R_FE: { /* nutrients.jl:22-23 */
            double t = p->nutrients == OBM_NUT_NITRATE_AMMONIA_IRON ? -nutrient_uptake_Fe(p, c) : 0.0;
            ORC_TERMS(t);
            return t;
        }
        case R_NO3: /* nutrients.jl:31-34 */
            if (!has_NA(p)) return 0.0;
            ORC_TERMS(nitrification(p, c), nutrient_uptake_NO3(p, c));
            return (nitrification(p, c) - nutrient_uptake_NO3(p, c));
        case R_NH4: /* nutrients.jl:36-41 */
            if (!has_NA(p)) return 0.0;
            {
                (void)nutrient_uptake_NH4(p, c);
                double s_up = orc_term_scale; /* uptake − exuded ammonia: its own Σ|terms| */
                ORC_TERMS(plankton_inorganic_nitrogen_waste(p, c), detritus_inorganic_nitrogen_waste(p, c), nitrification(p, c), s_up);
            }
            {
                double s_all = orc_term_scale; /* the uptake call below records its own scale again */
                double t = (plankton_inorganic_nitrogen_waste(p, c) + detritus_inorganic_nitrogen_waste(p, c)
                            - nitrification(p, c) - nutrient_uptake_NH4(p, c));
                orc_term_scale = s_all;
                return t;
            }
        case R_N: /* nutrients.jl:60-64 */
            if (has_NA(p)) return 0.0;
            ORC_TERMS(plankton_inorganic_nitrogen_waste(p, c), detritus_inorganic_nitrogen_waste(p, c), nutrient_uptake_N(p, c));
            return (plankton_inorganic_nitrogen_waste(p, c) + detritus_inorganic_nitrogen_waste(p, c)
                    - nutrient_uptake_N(p, c));
        case R_P: { /* plankton.jl:92-105 */
            double gamma = p->phytoplankton_exudation_fraction, m = p->phytoplankton_mortality_rate;
            double P = c->v[R_P];
            double muP = phytoplankton_growth(p, c);
            double Gp = grazing_P(p, c);
            double nu = mortality(p->phytoplankton_mortality_formulation, P, m);
            ORC_TERMS((1 - gamma) * muP, Gp, nu);
            return (1 - gamma) * muP - Gp - nu;
        }
        case R_Z: { /* plankton.jl:107-116 */
            double a = p->zooplankton_assimilation_fraction, m = p->zooplankton_mortality_rate, mu = p->zooplankton_excretion_rate;
            double Z = c->v[R_Z];
            double G = total_grazing(p, c);
            ORC_TERMS(a * G, m * (Z * Z), mu * Z);
            return a * G - m * (Z * Z) - mu * Z;
        }
        case R_D: /* detritus.jl:282-288 */
            if (p->detritus != OBM_DET_DETRITUS) return 0.0;
            ORC_TERMS(plankton_organic_nitrogen_waste(p, c), solid_waste(p, c), grazing_sPOM(p, c), p->remineralisation_rate * c->v[R_D]);
            return (plankton_organic_nitrogen_waste(p, c) + solid_waste(p, c) - grazing_sPOM(p, c)
                    - p->remineralisation_rate * c->v[R_D]);
        case R_SPOM: /* detritus.jl:149-153 */
            ORC_TERMS(p->small_solid_waste_fraction * solid_waste(p, c), grazing_sPOM(p, c),
                      p->small_remineralisation_rate * small_particulate_concentration(p, c));
            return (p->small_solid_waste_fraction * solid_waste(p, c) - grazing_sPOM(p, c)
                    - p->small_remineralisation_rate * small_particulate_concentration(p, c));
        case R_BPOM: /* detritus.jl:155-158 */
            ORC_TERMS((1 - p->small_solid_waste_fraction) * solid_waste(p, c), p->large_remineralisation_rate * large_particulate_concentration(p, c));
            return ((1 - p->small_solid_waste_fraction) * solid_waste(p, c)
                    - p->large_remineralisation_rate * large_particulate_concentration(p, c));
        case R_DOM: /* detritus.jl:143-147 */
            ORC_TERMS(plankton_organic_nitrogen_waste(p, c), detritus_organic_nitrogen_waste(p, c),
                      p->dissolved_remineralisation_rate * dissolved_organic_nitrogen(p, c));
            return (plankton_organic_nitrogen_waste(p, c) + detritus_organic_nitrogen_waste(p, c)
                    - p->dissolved_remineralisation_rate * dissolved_organic_nitrogen(p, c));
        case R_SPOC: /* detritus.jl:85-89 */
            ORC_TERMS(p->small_solid_waste_fraction * solid_carbon_waste(p, c), grazing_sPOC(p, c),
                      p->small_remineralisation_rate * small_particulate_carbon_concentration(p, c));
            return (p->small_solid_waste_fraction * solid_carbon_waste(p, c) - grazing_sPOC(p, c)
                    - p->small_remineralisation_rate * small_particulate_carbon_concentration(p, c));
        case R_BPOC: /* detritus.jl:91-95 */
            ORC_TERMS((1 - p->small_solid_waste_fraction) * solid_carbon_waste(p, c), calcite_production(p, c),
                      p->large_remineralisation_rate * large_particulate_carbon_concentration(p, c));
            return ((1 - p->small_solid_waste_fraction) * solid_carbon_waste(p, c) + calcite_production(p, c)
                    - p->large_remineralisation_rate * large_particulate_carbon_concentration(p, c));
        case R_DOC: /* detritus.jl:97-101 */
            ORC_TERMS(plankton_organic_carbon_waste(p, c), detritus_organic_carbon_waste(p, c),
                      p->dissolved_remineralisation_rate * dissolved_organic_carbon(p, c));
            return (plankton_organic_carbon_waste(p, c) + detritus_organic_carbon_waste(p, c)
                    - p->dissolved_remineralisation_rate * dissolved_organic_carbon(p, c));
        case R_DIC: /* carbonate_system.jl:50-55 */
            ORC_TERMS(phytoplankton_primary_production(p, c), plankton_inorganic_carbon_waste(p, c),
                      detritus_inorganic_carbon_waste(p, c), calcite_dissolution(p, c));
            return (-phytoplankton_primary_production(p, c) + plankton_inorganic_carbon_waste(p, c)
                    + detritus_inorganic_carbon_waste(p, c) + calcite_dissolution(p, c));
        case R_ALK: /* carbonate_system.jl:57-68 — a sum of tendencies: the un-cancelled scale is the sum of theirs */
            if (has_NA(p)) {
                (void)tendency(p, c, R_NH4);
                double s_nh4 = orc_term_scale;
                (void)tendency(p, c, R_NO3);
                double s_no3 = orc_term_scale;
                ORC_TERMS(s_nh4 * (1 - 1.0 / 16), s_no3 * (1 + 1.0 / 16), 2.0 * calcite_uptake(p, c), 2.0 * calcite_dissolution(p, c));
            } else {
                (void)tendency(p, c, R_N);
                double s_n = orc_term_scale;
                ORC_TERMS(s_n, 2.0 * calcite_uptake(p, c), 2.0 * calcite_dissolution(p, c));
            }
            {
                double s_all = orc_term_scale; /* the tendency calls below record their own scales again */
                double t;
                if (has_NA(p))
                    t = (tendency(p, c, R_NH4) * (1 - 1.0 / 16) - tendency(p, c, R_NO3) * (1 + 1.0 / 16)
                         - 2.0 * calcite_uptake(p, c) + 2.0 * calcite_dissolution(p, c));
                else
                    t = (tendency(p, c, R_N) - 2.0 * calcite_uptake(p, c) + 2.0 * calcite_dissolution(p, c));
                orc_term_scale = s_all;
                return t;
            }
        case R_O2: { /* oxygen.jl:21-31.  The `<:Nutrient` specialisation :33-43 dispatches on the
                        PLANKTON slot and is unreachable (SURVEY App. A bug 2), so Nutrient models use
                        the generic method with bgc(Val(:NH₄)) = 0 and nitrification = 0. */
            double Rp = p->respiration_oxygen_nitrogen_ratio, Rn = p->nitrification_oxygen_nitrogen_ratio;
            double muP = phytoplankton_growth(p, c);
            double nitrate_production = tendency(p, c, R_NH4);
            double s_nh4 = has_NA(p) ? orc_term_scale : 0.0;
            double muNH4 = nitrification(p, c);
            ORC_TERMS(Rp * muP, (Rp - Rn) * s_nh4, Rp * muNH4);
            return Rp * muP - (Rp - Rn) * nitrate_production - Rp * muNH4;
        }
        default: orc_term_scale = 0.0; return 0.0; /* T etc.: zero(grid) */
    }
}

static void load_cell(const obm_grid* g, const int* roles, int nt, const double* const* tracers, const double* PAR,
                      int i, int j, int k, cell_t* c) {
    int64_t idx = cell_index(g, i, j, k);
    for (int r = 0; r < R_COUNT; r++) c->v[r] = 0.0;
    for (int n = nt - 1; n >= 0; n--) /* first replicate of DIC/Alk wins (they are never read anyway) */
        c->v[roles[n]] = tracers[n][idx];
    c->PAR = PAR[idx];
}

/* One pass = one tracer (one compute_Gc! launch of the reference). */
int orc_npd_tendency(const obm_grid* g, const obm_npd_params* p, const double* const* tracers, const double* PAR,
                     int tracer_index, double* G, int accumulate) {
    int roles[OBM_NPD_MAX_TRACERS];
    int nt = orc_npd_layout(p, roles, NULL);
    if (nt < 0) return nt;
    if (tracer_index < 0 || tracer_index >= nt) return OBM_ESIZE;
    int role = roles[tracer_index];
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < g->Nz; k++)
        for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
                cell_t c;
                load_cell(g, roles, nt, tracers, PAR, i, j, k, &c);
                double t = tendency(p, &c, role);
                int64_t idx = cell_index(g, i, j, k);
                if (accumulate) G[idx] += t; else G[idx] = t;
            }
    return 0;
}

/* Σ|additive terms| of every tendency (the parity metric's S): S[n] parent arrays in layout order, NULL skips */
int orc_npd_tendency_scales(const obm_grid* g, const obm_npd_params* p, const double* const* tracers, const double* PAR,
                            double* const* S) {
    int roles[OBM_NPD_MAX_TRACERS];
    int nt = orc_npd_layout(p, roles, NULL);
    if (nt < 0) return nt;
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < g->Nz; k++)
        for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
                cell_t c;
                load_cell(g, roles, nt, tracers, PAR, i, j, k, &c);
                int64_t idx = cell_index(g, i, j, k);
                for (int n = 0; n < nt; n++) {
                    if (!S[n]) continue;
                    orc_term_scale = 0.0;
                    (void)tendency(p, &c, roles[n]);
                    S[n][idx] = orc_term_scale;
                }
            }
    return 0;
}

/* All tracers, reference launch structure: one full-grid pass per tracer. */
int orc_npd_tendencies(const obm_grid* g, const obm_npd_params* p, const double* const* tracers, const double* PAR,
                       double* const* G, int accumulate) {
    int nt = orc_npd_layout(p, NULL, NULL);
    if (nt < 0) return nt;
    for (int n = 0; n < nt; n++) {
        if (!G[n]) continue;
        int rc = orc_npd_tendency(g, p, tracers, PAR, n, G[n], accumulate);
        if (rc) return rc;
    }
    return 0;
}
