/*
 * oracle_pisces.c — ORACLE (test infrastructure, not product): CPU restatement of the 24 PISCES
 * tracer tendencies of OceanBioME.jl v0.17.6, function by function, in the reference's operation
 * order and with the reference's LAUNCH STRUCTURE (one pass per tracer; every pass re-evaluates
 * nutrient limitation, growth rates, grazing … exactly as the per-tracer callables do).
 *
 * Follows src/Models/AdvectedPopulations/PISCES/ :
 *   phytoplankton/{nutrient_limitation,growth_rate,mixed_mondo,mixed_mondo_nano_diatoms,nano_and_diatoms}.jl
 *   zooplankton/{food_quality_dependant,grazing_waste,iron_grazing,mortality_waste,micro_and_meso,defaults}.jl
 *   dissolved_organic_matter/dissolved_organic_carbon.jl
 *   particulate_organic_matter/{two_size_class,carbon,iron,silicate,calcite,nano_diatom_coupling,micro_meso_zoo_coupling}.jl
 *   nitrogen/nitrate_ammonia.jl, iron/{iron,simple_iron}.jl, silicate.jl, oxygen.jl, phosphate.jl,
 *   inorganic_carbon.jl, common.jl                      (cited per function as file:line)
 * including the reference's quirks: swapped day-length arguments, (P, D, POC, Z)
 * prey ordering, Alk = NH₄ − NO₃ − 2 CaCO₃, negative "sinking flux" in flux feeding.
 *
 * Pinned by the reference's own conservation test (test/test_PISCES.jl:32-92: C, Fe, Si, P, N
 * budgets of the tendencies close to atol 1e-20…1e-22 at PISCES_INITIAL_VALUES), the zero-state
 * test, and cross-checked against the independently derived smoke values of SURVEY §8c
 * (tests/test_oracle_pisces.py).  No absolute tendency goldens exist in the reference.
 */
#include "oracle_common.h"

enum { /* OBM_PISCES_NTRACERS order — PISCES.jl:94-105 */
    T_P = 0, T_PChl, T_PFe, T_D, T_DChl, T_DFe, T_DSi, T_Z, T_M, T_DOC, T_POC, T_GOC, T_SFe, T_BFe, T_PSi, T_CaCO3,
    T_NO3, T_NH4, T_PO4, T_Fe, T_Si, T_DIC, T_Alk, T_O2, T_T, T_S
};

typedef struct {
    double P, PChl, PFe, D, DChl, DFe, DSi, Z, M, DOC, POC, GOC, SFe, BFe, PSi, CaCO3, NO3, NH4, PO4, Fe, Si, DIC, Alk, O2, T, S;
    double PAR1, PAR2, PAR3, PAR, Omega;
    double wPOC, wGOC; /* ℑzᵃᵃᶜ(i, j, k, grid, w) = (w[k] + w[k+1]) / 2 */
    double zmxl, zeu, kappa, mlPAR;
    double z;
} pcell;

typedef const obm_pisces_params* PP;
typedef const obm_pisces_phyto* PH;
typedef const obm_pisces_zoo* ZO;
typedef const pcell* PC;

#define DAY 86400.0

__thread double orc_term_scale = 0.0; /* oracle_common.h: Σ|terms| of the tendency evaluated last */

static double jl_min3(double a, double b, double c) { return jl_min(jl_min(a, b), c); }
static double jl_min4(double a, double b, double c, double d) { return jl_min(jl_min(jl_min(a, b), c), d); }

/* common.jl:61-66 */
static double anoxia_factor(PP p, double O2) {
    return jl_min(1, jl_max(0, 0.4 * (p->first_anoxia_threshold - O2) / (p->second_anoxia_threshold + O2)));
}

/* ===================== phytoplankton ===================== */
typedef struct { double L, LFe, LPO4, LN, LNO3, LNH4; } nlim_t;

/* mixed_mondo.jl:207-215 */
static double size_factor(PH ph, double I) {
    double Im = ph->threshold_for_size_dependency, S = ph->size_ratio;
    double I1 = jl_min(I, Im);
    double I2 = jl_max(0, I - Im);
    return (I1 + S * I2) / (I1 + I2 + EPS0);
}
/* nutrient_limitation.jl:73 */
static double nitrogen_limitation(double N1, double N2, double K1, double K2) {
    return (K2 * N1) / (K1 * K2 + K1 * N2 + K2 * N1 + EPS0);
}
/* nutrient_limitation.jl:20-71 */
static nlim_t nutrient_limitation(PP p, PH ph, PC c, double I, double IChl, double IFe) {
    double kno = ph->minimum_nitrate_half_saturation, knh = ph->minimum_ammonium_half_saturation;
    double kp = ph->minimum_phosphate_half_saturation, ksi = ph->minimum_silicate_half_saturation;
    double pk = ph->silicate_half_saturation_parameter, theta_o = ph->optimal_iron_quota;
    double Sip = p->silicate_climatology;
    double tFe = I == 0 ? 0 : IFe / (I + EPS0);
    double tChl = I == 0 ? 0 : IChl / (12 * I + EPS0);
    double Kbar = size_factor(ph, I);
    double Kno = kno * Kbar, Knh = knh * Kbar, Kp = kp * Kbar, Ksi = ksi * Kbar;
    nlim_t r;
    r.LNO3 = nitrogen_limitation(c->NO3, c->NH4, Kno, Knh);
    r.LNH4 = nitrogen_limitation(c->NH4, c->NO3, Knh, Kno);
    r.LN = r.LNO3 + r.LNH4;
    r.LPO4 = c->PO4 / (c->PO4 + Kp + EPS0);
    double tm = 1000 * (0.0016 / 55.85 * 12 * tChl + 1.5 * 1.21e-5 * 14 / (55.85 * 7.625) * r.LN
                        + 1.15e-4 * 14 / (55.85 * 7.625) * r.LNO3);
    r.LFe = jl_min(1, jl_max(0, (tFe - tm) / theta_o));
    double KSi = Ksi + 7 * (Sip * Sip) / (pk * pk + Sip * Sip);
    double LSi = c->Si / (c->Si + KSi);
    LSi = ph->silicate_limited ? LSi : INFINITY;
#ifdef ORC_SELECT_MINMAX /* the fast pass's operand order: a select propagates a NaN in its SECOND operand, and L_Fe is the
                          * only limitation a finite state can turn into NaN (Inf − Inf of overflowed quotas) */
    r.L = jl_min4(LSi, r.LN, r.LPO4, r.LFe);
#else
    r.L = jl_min4(r.LN, r.LPO4, r.LFe, LSi);
#endif
    return r;
}

/* growth_rate.jl:158-165 */
static double base_production_rate(PH ph, double T) { return ph->base_growth_rate * pow(ph->temperature_sensitivity, T); }

/* growth_rate.jl:77-83, 117-124 */
static double light_limitation(PH ph, double I, double IChl, double T, double PAR, double day_length, double L, double alpha) {
    double theta = IChl / (12 * I + EPS0);
    if (ph->growth_rate_kind == OBM_GROWTH_NUTRIENT_LIMITED) {
        double mui = base_production_rate(ph, T);
        return 1 - exp(-alpha * theta * PAR / (day_length * mui * L + EPS0));
    }
    double br = ph->basal_respiration_rate, mur = ph->reference_growth_rate;
    return 1 - exp(-alpha * theta * PAR / (day_length * (br + mur)));
}

/* growth_rate.jl:3-47 — (μ::BaseProduction)(…, L); day length with SWAPPED arguments (:30) */
static double growth_rate(PP p, PH ph, PC c, double I, double IChl, double L) {
    double PAR = ph->blue_light_absorption * c->PAR1 + ph->green_light_absorption * c->PAR2 + ph->red_light_absorption * c->PAR3;
    double day_length = p->day_length_growth;
    double d = jl_max(0, c->zeu - c->zmxl);
    double dark_residence_time = d * d / c->kappa;
    double ft = pow(ph->temperature_sensitivity, c->T);
    double mui = ph->base_growth_rate * ft;
    double f1 = 1.5 * day_length / (day_length + 0.5 * DAY);
    double f2 = 1 - dark_residence_time / (dark_residence_time + ph->dark_tolerance);
    double alpha = ph->initial_slope_of_PI_curve * (1 + ph->low_light_adaptation * exp(-PAR));
    double fl = light_limitation(ph, I, IChl, c->T, PAR, day_length, L, alpha);
    return mui * f1 * f2 * fl * L;
}

/* growth_rate.jl:126-156 — day length with the CORRECT argument order (:143) */
static void production_and_energy_assimilation_absorption_ratio(PP p, PH ph, PC c, double I, double IChl, double IFe,
                                                                double* mu_out, double* rho_out) {
    double PAR = ph->blue_light_absorption * c->PAR1 + ph->green_light_absorption * c->PAR2 + ph->red_light_absorption * c->PAR3;
    double day_length = p->day_length_chlorophyll;
    double f1 = 1.5 * day_length / (day_length + 0.5 * DAY);
    double L = nutrient_limitation(p, ph, c, I, IChl, IFe).L;
    double mu = growth_rate(p, ph, c, I, IChl, L);
    double mucheck = mu / f1 * day_length;
    double alpha = ph->initial_slope_of_PI_curve * (1 + ph->low_light_adaptation * exp(-PAR));
    *mu_out = mu;
    *rho_out = 12 * mucheck * I / (alpha * IChl * PAR + EPS0) * L;
}

/* mixed_mondo.jl:137-167 */
static void phyto_mortality(PP p, PH ph, PC c, double I, double IChl, double IFe, double* lin, double* quad) {
    double L = nutrient_limitation(p, ph, c, I, IChl, IFe).L;
    double K = ph->mortality_half_saturation, m = ph->linear_mortality_rate;
    *lin = m * I / (I + K) * I;
    double w0 = ph->base_quadratic_mortality, w1 = ph->maximum_quadratic_mortality;
    double w = w0 + w1 * 0.25 * (1 - L * L) / (0.25 + L * L);
    double shear = c->z < c->zmxl ? p->background_shear : p->mixed_layer_shear;
    *quad = shear * w * (I * I);
}

/* mixed_mondo.jl:169-175 */
static double total_production(PP p, PH ph, PC c, double I, double IChl, double IFe) {
    double L = nutrient_limitation(p, ph, c, I, IChl, IFe).L;
    return growth_rate(p, ph, c, I, IChl, L) * I;
}

/* mixed_mondo.jl:177-205 */
static double iron_uptake(PP p, PH ph, PC c, double I, double IChl, double IFe) {
    double delta = ph->exudated_fraction, tFem = ph->maximum_iron_ratio;
    double tFe = IFe / (I + EPS0);
    nlim_t nl = nutrient_limitation(p, ph, c, I, IChl, IFe);
    double mui = base_production_rate(ph, c->T);
    double K = ph->half_saturation_for_iron_uptake * size_factor(ph, I);
    double L1 = c->Fe / (c->Fe + K + EPS0);
    double L2 = 4 - 4.5 * nl.LFe / (nl.LFe + 1);
    return (1 - delta) * tFem * L1 * L2 * jl_max(0, (1 - tFe / tFem) / (1.05 - tFe / tFem)) * mui * I;
}

/* mixed_mondo.jl:217-248 */
static double silicate_uptake(PP p, PH ph, PC c, double I, double IChl, double IFe) {
    double delta = ph->exudated_fraction, K1 = ph->silicate_half_saturation, K2 = ph->enhanced_silicate_half_saturation;
    double t0 = ph->optimal_silicate_ratio;
    double Si = c->Si;
    nlim_t nl = nutrient_limitation(p, ph, c, I, IChl, IFe);
    double mu = growth_rate(p, ph, c, I, IChl, nl.L);
    double mui = base_production_rate(ph, c->T);
    double L1 = Si / (Si + K1 + EPS0);
    double L2 = p->latitude < 0 ? (Si * Si * Si) / (Si * Si * Si + K2 * K2 * K2) : 0;
    double F1 = jl_min4(mu / (mui * nl.L + EPS0), nl.LFe, nl.LPO4, nl.LN);
    double F2 = jl_min(1, 2.2 * jl_max(0, L1 - 0.5));
    double t1 = t0 * L1 * jl_min(5.4, (4.4 * exp(-4.23 * F1) * F2 + 1) * (1 + 2 * L2));
    return (1 - delta) * t1 * mu * I;
}

/* mixed_mondo.jl:250-267 */
static double uptake_NO3(PP p, PH ph, PC c, double I, double IChl, double IFe) {
    nlim_t nl = nutrient_limitation(p, ph, c, I, IChl, IFe);
    double muI = total_production(p, ph, c, I, IChl, IFe);
    return muI * nl.LNO3 / (nl.LN + EPS0);
}
static double uptake_NH4(PP p, PH ph, PC c, double I, double IChl, double IFe) {
    nlim_t nl = nutrient_limitation(p, ph, c, I, IChl, IFe);
    double muI = total_production(p, ph, c, I, IChl, IFe);
    return muI * nl.LNH4 / (nl.LN + EPS0);
}

#define NANO &p->nano, c, c->P, c->PChl, c->PFe
#define DIAT &p->diatoms, c, c->D, c->DChl, c->DFe

/* ===================== zooplankton ===================== */
/* flux_rate two_size_class.jl:95-98 */
static double flux_POC(PC c) { return c->POC * c->wPOC; }
static double flux_GOC(PC c) { return c->GOC * c->wGOC; }
static double flux_SFe(PC c) { return c->SFe * c->wPOC; }
static double flux_BFe(PC c) { return c->BFe * c->wGOC; }

typedef struct { double tsg, avail, ge; double food[4], iron[4]; int N; } graze_t;

/* the common front half of food_quality_dependant.jl:126-169 / :226-255 / iron_grazing.jl:2-34.
 * Prey NamedTuple is (P, D, POC, Z) for every class (defaults.jl:45-60); the sums run over the
 * first N = length(prey_names) entries: 3 for Z, 4 for M. */
static graze_t grazing_core(PP p, ZO zoo, PC c, int is_meso) {
    graze_t g;
    g.N = is_meso ? 4 : 3;
    double J = zoo->specific_food_threshold_concentration, K = zoo->grazing_half_saturation;
    double base = zoo->maximum_grazing_rate * pow(zoo->temperature_sensitivity, c->T);
    g.food[0] = c->P; g.food[1] = c->D; g.food[2] = c->POC; g.food[3] = c->Z;
    double total_food = 0, avail = 0;
    for (int n = 0; n < g.N; n++) total_food = n == 0 ? g.food[n] * zoo->food_preferences[n] : total_food + g.food[n] * zoo->food_preferences[n];
    for (int n = 0; n < g.N; n++) {
        double t = jl_max(0, (g.food[n] - J)) * zoo->food_preferences[n];
        avail = n == 0 ? t : avail + t;
    }
    double clg = jl_max(0, avail - jl_min(avail / 2, zoo->food_threshold_concentration));
    g.tsg = base * clg / (K + total_food);
    g.avail = avail;
    g.iron[0] = c->PFe / (c->P + EPS0);
    g.iron[1] = c->DFe / (c->D + EPS0);
    g.iron[2] = c->SFe / (c->POC + EPS0);
    g.iron[3] = p->micro.iron_ratio;
    double total_iron = 0;
    for (int n = 0; n < g.N; n++) total_iron = n == 0 ? g.iron[n] * zoo->food_preferences[n] : total_iron + g.iron[n] * zoo->food_preferences[n];
    double igr = total_iron / (zoo->iron_ratio * g.tsg + EPS0);
    double fq = jl_min(1, igr);
    g.ge = fq * jl_min(zoo->minimum_growth_efficiency, (1 - zoo->non_assimilated_fraction) * igr);
    return g;
}
static double zoo_I(PC c, int is_meso) { return is_meso ? c->M : c->Z; }

/* food_quality_dependant.jl:126-169 → (gI, e) */
static double zoo_grazing(PP p, ZO zoo, PC c, int is_meso, double* e) {
    graze_t g = grazing_core(p, zoo, c, is_meso);
    if (e) *e = g.ge;
    return g.tsg * zoo_I(c, is_meso);
}
/* :171-186 */
static double zoo_flux_feeding(PP p, ZO zoo, PC c, int is_meso) {
    (void)p;
    double sinking_flux = flux_POC(c) + flux_GOC(c);
    double base = zoo->maximum_flux_feeding_rate * pow(zoo->temperature_sensitivity, c->T);
    double tsff = base * sinking_flux;
    return tsff * zoo_I(c, is_meso);
}
/* :188-204 */
static double zoo_mortality(PP p, ZO zoo, PC c, int is_meso) {
    double I = zoo_I(c, is_meso);
    double tf = pow(zoo->temperature_sensitivity, c->T);
    double cf = I / (I + zoo->mortality_half_saturation);
    return tf * I * (zoo->quadratic_mortality * I + zoo->linear_mortality * (cf + 3 * anoxia_factor(p, c->O2)));
}
/* :206-220 */
static double zoo_linear_mortality(PP p, ZO zoo, PC c, int is_meso) {
    double I = zoo_I(c, is_meso);
    double tf = pow(zoo->temperature_sensitivity, c->T);
    double cf = I / (I + zoo->mortality_half_saturation);
    return tf * zoo->linear_mortality * (cf + 3 * anoxia_factor(p, c->O2)) * I;
}
/* :226-255 — grazing on one prey; prey index into (P, D, POC, Z), or -1 (no preference ⇒ 0) */
static double zoo_grazing_on(PP p, ZO zoo, PC c, int is_meso, int prey) {
    graze_t g = grazing_core(p, zoo, c, is_meso);
    double pref = prey < 0 ? 0 : zoo->food_preferences[prey];
    double Pc = prey < 0 ? 0 : g.food[prey];
    double J = zoo->specific_food_threshold_concentration;
    return pref * jl_max(0, Pc - J) * g.tsg / (g.avail + EPS0) * zoo_I(c, is_meso);
}
/* :257-272 */
static double zoo_flux_feeding_on(ZO zoo, PC c, int is_meso, double flux) {
    double base = zoo->maximum_flux_feeding_rate * pow(zoo->temperature_sensitivity, c->T);
    return base * flux * zoo_I(c, is_meso);
}
/* :111-117 */
static double growth_death(PP p, ZO zoo, PC c, int is_meso) {
    double e;
    double gI = zoo_grazing(p, zoo, c, is_meso, &e);
    double gfI = zoo_flux_feeding(p, zoo, c, is_meso);
    double mI = zoo_mortality(p, zoo, c, is_meso);
    ORC_TERMS(e * gI, e * gfI, mI);
    return e * (gI + gfI) - mI;
}
/* micro_and_meso.jl:50-52: grazing(zoo::MicroAndMeso, prey) = micro + meso */
static double grazing_both(PP p, PC c, int prey) {
    return zoo_grazing_on(p, &p->micro, c, 0, prey) + zoo_grazing_on(p, &p->meso, c, 1, prey);
}
static double flux_feeding_both(PP p, PC c, double flux) {
    return zoo_flux_feeding_on(&p->micro, c, 0, flux) + zoo_flux_feeding_on(&p->meso, c, 1, flux);
}
/* grazing_waste.jl:3-33 */
static double non_assimilated_waste(PP p, ZO zoo, PC c, int is_meso) {
    double gI = zoo_grazing(p, zoo, c, is_meso, NULL);
    double gfI = zoo_flux_feeding(p, zoo, c, is_meso);
    return zoo->non_assimilated_fraction * (gI + gfI);
}
static double excretion(PP p, ZO zoo, PC c, int is_meso) {
    double e;
    double gI = zoo_grazing(p, zoo, c, is_meso, &e);
    double gfI = zoo_flux_feeding(p, zoo, c, is_meso);
    return (1 - zoo->non_assimilated_fraction - e) * (gI + gfI);
}
static double inorganic_excretion1(PP p, ZO zoo, PC c, int m) { return zoo->dissolved_excretion_fraction * excretion(p, zoo, c, m); }
static double organic_excretion1(PP p, ZO zoo, PC c, int m) { return (1 - zoo->dissolved_excretion_fraction) * excretion(p, zoo, c, m); }
static double inorganic_excretion(PP p, PC c) { return inorganic_excretion1(p, &p->micro, c, 0) + inorganic_excretion1(p, &p->meso, c, 1); }
static double organic_excretion(PP p, PC c) { return organic_excretion1(p, &p->micro, c, 0) + organic_excretion1(p, &p->meso, c, 1); }

/* iron_grazing.jl:2-34 */
static double iron_grazing(PP p, ZO zoo, PC c, int is_meso) {
    graze_t g = grazing_core(p, zoo, c, is_meso);
    double J = zoo->specific_food_threshold_concentration;
    double s = 0;
    for (int n = 0; n < g.N; n++) {
        double t = jl_max(0, (g.food[n] - J)) * zoo->food_preferences[n] * g.iron[n];
        s = n == 0 ? t : s + t;
    }
    double tsig = s * g.tsg / (g.avail + EPS0);
    return tsig * zoo_I(c, is_meso);
}
/* iron_grazing.jl:36-51 */
static double iron_flux_feeding(ZO zoo, PC c, int is_meso) {
    double sinking_flux = flux_SFe(c) + flux_BFe(c);
    double base = zoo->maximum_flux_feeding_rate * pow(zoo->temperature_sensitivity, c->T);
    return base * sinking_flux * zoo_I(c, is_meso);
}
/* grazing_waste.jl:35-43 */
static double non_assimilated_iron_waste(PP p, ZO zoo, PC c, int is_meso) {
    double gI = iron_grazing(p, zoo, c, is_meso);
    double gfI = iron_flux_feeding(zoo, c, is_meso);
    return zoo->non_assimilated_fraction * (gI + gfI);
}
/* grazing_waste.jl:45-63 */
static double non_assimilated_iron1(PP p, ZO zoo, PC c, int is_meso) {
    double ge;
    double gI = zoo_grazing(p, zoo, c, is_meso, &ge);
    double gfI = zoo_flux_feeding(p, zoo, c, is_meso);
    double zoo_assimilated_iron = zoo->iron_ratio * ge * (gI + gfI);
    double gIFe = iron_grazing(p, zoo, c, is_meso);
    double gfIFe = iron_flux_feeding(zoo, c, is_meso);
    double lost_to_particles = zoo->non_assimilated_fraction * (gIFe + gfIFe);
    double total_iron_grazed = gIFe + gfIFe;
    ORC_TERMS(total_iron_grazed, lost_to_particles, zoo_assimilated_iron);
    return total_iron_grazed - lost_to_particles - zoo_assimilated_iron;
}
static double non_assimilated_iron(PP p, PC c) {
    double micro = non_assimilated_iron1(p, &p->micro, c, 0);
    double s_micro = orc_term_scale;
    double meso = non_assimilated_iron1(p, &p->meso, c, 1);
    orc_term_scale += s_micro;
    return micro + meso;
}
/* grazing_waste.jl:65-71; micro_and_meso.jl:134-136 */
static double calcite_loss(PP p, PC c, int prey) {
    return p->micro.undissolved_calcite_fraction * zoo_grazing_on(p, &p->micro, c, 0, prey)
           + p->meso.undissolved_calcite_fraction * zoo_grazing_on(p, &p->meso, c, 1, prey);
}
/* mortality_waste.jl:2-41 — all for the meso class (micro_and_meso.jl:70-83) */
static double upper_trophic_waste(PP p, PC c) {
    ZO zoo = &p->meso;
    double tf = pow(zoo->temperature_sensitivity, c->T);
    return 1 / (1 - zoo->minimum_growth_efficiency) * zoo->quadratic_mortality * tf * (c->M * c->M);
}
static double upper_trophic_respiration_product(PP p, PC c) {
    return (1 - p->meso.minimum_growth_efficiency - p->meso.non_assimilated_fraction) * upper_trophic_waste(p, c);
}
static double upper_trophic_excretion(PP p, PC c) { return (1 - p->meso.dissolved_excretion_fraction) * upper_trophic_respiration_product(p, c); }
static double upper_trophic_respiration(PP p, PC c) { return p->meso.dissolved_excretion_fraction * upper_trophic_respiration_product(p, c); }
static double upper_trophic_dissolved_iron(PP p, PC c) { return p->meso.iron_ratio * upper_trophic_respiration_product(p, c); }
static double upper_trophic_fecal_production(PP p, PC c) { return p->meso.non_assimilated_fraction * upper_trophic_waste(p, c); }
static double upper_trophic_fecal_iron_production(PP p, PC c) { return upper_trophic_fecal_production(p, c) * p->meso.iron_ratio; }

/* micro_and_meso.jl:85-105 */
static double bacteria_concentration(PP p, PC c) {
    double bZ = p->microzooplankton_bacteria_concentration, bM = p->mesozooplankton_bacteria_concentration;
    double a = p->bacteria_concentration_depth_exponent;
    double zm = jl_min(c->zmxl, c->zeu);
    double surface_bacteria = jl_min(4, bZ * c->Z + bM * c->M);
    double depth_factor = pow(zm / c->z, a);
    return (c->z >= zm ? 1 : depth_factor) * surface_bacteria;
}
/* micro_and_meso.jl:107-132 */
static double bacteria_activity(PP p, PC c) {
    double K_DOC = p->doc_half_saturation_for_bacterial_activity, K_NO3 = p->nitrate_half_saturation_for_bacterial_activity;
    double K_NH4 = p->ammonia_half_saturation_for_bacterial_activity, K_PO4 = p->phosphate_half_saturation_for_bacterial_activity;
    double K_Fe = p->iron_half_saturation_for_bacterial_activity;
    double DOC_limit = c->DOC / (c->DOC + K_DOC);
    double L_N = (K_NO3 * c->NH4 + K_NH4 * c->NO3) / (K_NO3 * K_NH4 + K_NO3 * c->NH4 + K_NH4 * c->NO3);
    double L_PO4 = c->PO4 / (c->PO4 + K_PO4);
    double L_Fe = c->Fe / (c->Fe + K_Fe);
    double limiting_quota = jl_min3(L_N, L_PO4, L_Fe);
    return limiting_quota * DOC_limit;
}

/* ===================== iron chemistry ===================== */
/* iron/iron.jl:25-37 */
/* Fe′ is the small root of a quadratic written as a difference: −Δ + √(Δ² + 4K·Fe) with Δ = 1 + K·L − K·Fe, itself a
 * difference.  orc_fep_scale records ITS Σ|additive terms| (Δ's included): the tendencies that multiply Fe′ use it in
 * place of |Fe′| when orc_nested_free_iron is set — the convention of oracle_common.h ("a term that is itself a
 * difference contributes its own Σ|terms|") applied one level further down, for states where Fe′ is pure cancellation
 * noise (ligands ≫ Fe or Fe ≫ ligands by ten orders: tests of extreme finite states only; off by default). */
int orc_nested_free_iron = 0;
static __thread double orc_fep_scale = 0.0;
void orc_pisces_set_nested_free_iron_scale(int on) { orc_nested_free_iron = on; }
static double free_iron(PC c) {
    double ligands = jl_max(0.6, 0.09 * (c->DOC + 40) - 3);
    double K = exp(16.27 - 1565.7 / jl_max(c->T + 273.15, 5));
    double D = 1 + K * ligands - K * c->Fe;
    double root = sqrt(D * D + 4 * K * c->Fe);
    orc_fep_scale = orc_nested_free_iron ? (1 + fabs(K * ligands) + fabs(K * c->Fe) + root) / fabs(2 * K) : fabs((-D + root) / (2 * K));
    return (-D + root) / (2 * K);
}

/* ===================== dissolved organic matter ===================== */
/* dissolved_organic_carbon.jl:56-71 */
static double dom_degradation(PP p, PC c) {
    double f = pow(p->dom_temperature_sensitivity, c->T);
    double Bact = bacteria_concentration(p, c);
    double LBact = bacteria_activity(p, c);
    return p->dom_remineralisation_rate * f * LBact * Bact / p->dom_reference_bacteria_concentration * c->DOC;
}
/* :73-94 → Φ₁, Φ₂, Φ₃ (total = Φ₁ + Φ₂ + Φ₃) */
static void dom_aggregation(PP p, PC c, double* total, double* F1, double* F2, double* F3) {
    const double* a = p->dom_aggregation_parameters;
    double shear = c->z < c->zmxl ? p->background_shear : p->mixed_layer_shear;
    double P1 = shear * (a[0] * c->DOC + a[1] * c->POC) * c->DOC;
    double P2 = shear * (a[2] * c->GOC) * c->DOC;
    double P3 = (a[3] * c->POC + a[4] * c->DOC) * c->DOC;
    if (total) *total = P1 + P2 + P3;
    if (F1) *F1 = P1;
    if (F2) *F2 = P2;
    if (F3) *F3 = P3;
}
/* :96-110 → (CgFe1 + CgFe2, CgFe1, CgFe2) */
static __thread double orc_cg_scale[2]; /* Σ|terms| of CgFe1, CgFe2 (see free_iron) */
static void aggregation_of_colloidal_iron(PP p, PC c, double* total, double* Cg1, double* Cg2) {
    double F1, F2, F3;
    dom_aggregation(p, c, NULL, &F1, &F2, &F3);
    double Fep = free_iron(c);
    double ligand_iron = c->Fe - Fep;
    double colloidal_iron = 0.5 * ligand_iron;
    double C1 = (F1 + F3) * colloidal_iron / (c->DOC + EPS0);
    double C2 = F2 * colloidal_iron / (c->DOC + EPS0);
    double li_s = orc_nested_free_iron ? fabs(c->Fe) + orc_fep_scale : fabs(ligand_iron); /* Σ|terms| of Fe − Fe′ */
    orc_cg_scale[0] = fabs((F1 + F3) * 0.5 * li_s / (c->DOC + EPS0));
    orc_cg_scale[1] = fabs(F2 * 0.5 * li_s / (c->DOC + EPS0));
    if (total) *total = C1 + C2;
    if (Cg1) *Cg1 = C1;
    if (Cg2) *Cg2 = C2;
}
/* :112-130 */
static double oxic_remineralisation(PP p, PC c) { return (1 - anoxia_factor(p, c->O2)) * dom_degradation(p, c); }
static double anoxic_remineralisation(PP p, PC c) { return anoxia_factor(p, c->O2) * dom_degradation(p, c); }

/* ===================== particulate organic matter ===================== */
/* two_size_class.jl:109-125 */
static double pom_aggregation(PP p, PC c) {
    const double* a = p->pom_aggregation_parameters;
    double shear = c->z < c->zmxl ? p->background_shear : p->mixed_layer_shear;
    return shear * (a[0] * (c->POC * c->POC) + a[1] * c->POC * c->GOC) + a[2] * c->POC * c->GOC + a[3] * (c->POC * c->POC);
}
/* :127-137 */
static double specific_degradation_rate(PP p, PC c) {
    double dO2 = anoxia_factor(p, c->O2);
    return p->pom_base_breakdown_rate * pow(p->pom_temperature_sensitivity, c->T) * (1 - 0.45 * dO2);
}
/* iron.jl:97-107 */
static double iron_scavenging_rate(PP p, PC c) {
    return p->minimum_iron_scavenging_rate + p->load_specific_iron_scavenging_rate * (c->POC + c->GOC + c->CaCO3 + c->PSi);
}
/* iron.jl:109-126 */
static double bacterial_iron_uptake(PP p, PC c) {
    double mu = p->maximum_bacterial_growth_rate * pow(p->pom_temperature_sensitivity, c->T);
    double Bact = bacteria_concentration(p, c);
    double LBact = bacteria_activity(p, c);
    return mu * LBact * p->maximum_iron_ratio_in_bacteria * c->Fe / (c->Fe + p->iron_half_saturation_for_bacteria) * Bact
           * p->bacterial_iron_uptake_efficiency;
}
/* iron.jl:128-137 */
static double iron_scavenging(PP p, PC c) { return iron_scavenging_rate(p, c) * (c->POC + c->GOC) * free_iron(c); }

/* nano_diatom_coupling.jl:57-71 */
static double coccolithophore_nutrient_limitation(PP p, PC c) {
    nlim_t nl = nutrient_limitation(p, NANO);
    double L_Fe = c->Fe / (c->Fe + 0.05);
    return jl_min3(nl.LN, L_Fe, nl.LPO4);
}
/* nano_diatom_coupling.jl:96-124 */
static double rain_ratio(PP p, PC c) {
    double L_CaCO3 = coccolithophore_nutrient_limitation(p, c);
    double pcf = jl_max(1, c->P / 2);
    double low_light_factor = jl_max(0, c->PAR - 1) / (4 + c->PAR);
    double high_light_factor = 30 / (30 + c->PAR);
    double low_temperature_factor = jl_max(0, c->T / (c->T + 0.1));
    double high_temperature_factor = 1 + exp(-((c->T - 10) * (c->T - 10)) / 25);
    double depth_factor = jl_min(1, -50 / c->zmxl);
    return (p->base_rain_ratio * L_CaCO3 * pcf * low_light_factor * high_light_factor * low_temperature_factor
            * high_temperature_factor * depth_factor);
}
/* nano_diatom_coupling.jl:1-55 */
static double small_mortality_phyto(PP p, PC c) {
    double Pl, Pq, Dl, Dq;
    phyto_mortality(p, NANO, &Pl, &Pq);
    double R = rain_ratio(p, c);
    phyto_mortality(p, DIAT, &Dl, &Dq);
    return (1 - R / 2) * (Pl + Pq) + Dl / 2;
}
static double large_mortality_phyto(PP p, PC c) {
    double Pl, Pq, Dl, Dq;
    phyto_mortality(p, NANO, &Pl, &Pq);
    double R = rain_ratio(p, c);
    phyto_mortality(p, DIAT, &Dl, &Dq);
    return R / 2 * (Pl + Pq) + Dl / 2 + Dq;
}
static double small_mortality_iron_phyto(PP p, PC c) {
    double Pl, Pq, Dl, Dq;
    phyto_mortality(p, NANO, &Pl, &Pq);
    double R = rain_ratio(p, c);
    phyto_mortality(p, DIAT, &Dl, &Dq);
    double tP = c->PFe / (c->P + EPS0), tD = c->DFe / (c->D + EPS0);
    return (1 - R / 2) * (Pl + Pq) * tP + Dl * tD / 2;
}
static double large_mortality_iron_phyto(PP p, PC c) {
    double Pl, Pq, Dl, Dq;
    phyto_mortality(p, NANO, &Pl, &Pq);
    double R = rain_ratio(p, c);
    phyto_mortality(p, DIAT, &Dl, &Dq);
    double tP = c->PFe / (c->P + EPS0), tD = c->DFe / (c->D + EPS0);
    return R / 2 * (Pl + Pq) * tP + (Dl / 2 + Dq) * tD;
}
/* nano_diatom_coupling.jl:73-84 */
static double particulate_silicate_production(PP p, PC c) {
    double theta = c->DSi / (c->D + EPS0);
    double Dl, Dq;
    phyto_mortality(p, DIAT, &Dl, &Dq);
    double tg = grazing_both(p, c, 1);
    return (tg + Dl + Dq) * theta;
}
/* nano_diatom_coupling.jl:86-94 */
static double calcite_production(PP p, PC c) {
    double R = rain_ratio(p, c);
    double l, q;
    phyto_mortality(p, NANO, &l, &q);
    double tgl = calcite_loss(p, c, 0);
    return R * (tgl + (l + q) / 2);
}
/* calcite.jl:9-19 */
static double calcite_dissolution(PP p, PC c) {
    double dCa = jl_max(0, 1 - c->Omega);
    return p->base_calcite_dissolution_rate * pow(dCa, p->calcite_dissolution_exponent) * c->CaCO3;
}
/* silicate.jl:32-48 */
static double particulate_silicate_liable_fraction(PP p, PC c) {
    double ll = p->fast_dissolution_rate_of_silicate, lr = p->slow_dissolution_rate_of_silicate;
    double zm = jl_min(c->zmxl, c->zeu);
    return p->base_liable_silicate_fraction * (c->z >= zm ? 1 : exp((ll - lr) * (zm - c->z) / c->wGOC));
}
/* silicate.jl:11-30 */
static double particulate_silicate_dissolution(PP p, PC c) {
    double ll = p->fast_dissolution_rate_of_silicate, lr = p->slow_dissolution_rate_of_silicate;
    double chi = particulate_silicate_liable_fraction(p, c);
    double l0 = chi * ll + (1 - chi) * lr;
    double equilibrium_silicate = pow(10.0, 6.44 - 968 / (c->T + 273.15));
    double silicate_saturation = (equilibrium_silicate - c->Si) / equilibrium_silicate;
    double l = l0 * (0.225 * (1 + c->T / 15) * silicate_saturation
                     + 0.775 * pow(pow(1 + c->T / 400, 4.0) * silicate_saturation, 9.0));
    return l * c->PSi;
}
/* micro_meso_zoo_coupling.jl:27-32 */
static double total_grazing_POC(PP p, PC c) { return grazing_both(p, c, 2) + flux_feeding_both(p, c, flux_POC(c)); }
static double total_grazing_GOC(PP p, PC c) { return flux_feeding_both(p, c, flux_GOC(c)); }

/* ===================== nitrogen ===================== */
/* nitrate_ammonia.jl:52-63 */
static double nitrification(PP p, PC c) {
    return p->maximum_nitrification_rate * c->NH4 / (1 + c->mlPAR) * (1 - anoxia_factor(p, c->O2));
}
/* nitrate_ammonia.jl:65-89 */
static double nitrogen_fixation(PP p, PC c) {
    double availability_limitation = nutrient_limitation(p, NANO).LN;
    double fixation_limit = availability_limitation >= 0.8 ? 0.01 : 1 - availability_limitation;
    double mu = base_production_rate(&p->nano, c->T);
    double growth_requirement = jl_max(0, mu - 2.15);
    double nutrient_lim = jl_min(c->Fe / (c->Fe + p->iron_half_saturation_for_fixation),
                                 c->PO4 / (c->PO4 + p->phosphate_half_saturation_for_fixation));
    double light_lim = 1 - exp(-c->PAR / p->light_saturation_for_fixation);
    return p->maximum_fixation_rate * growth_requirement * fixation_limit * nutrient_lim * light_lim;
}

/* ===================== the 24 per-tracer callables ===================== */
static double tendency(PP p, PC c, int name);

static double phyto_carbon(PP p, PH ph, PC c, double I, double IChl, double IFe, int prey) { /* mixed_mondo_nano_diatoms.jl:45-57 */
    double growth = (1 - ph->exudated_fraction) * total_production(p, ph, c, I, IChl, IFe);
    double l, q;
    phyto_mortality(p, ph, c, I, IChl, IFe, &l, &q);
    double death = (l + q);
    double grazed = grazing_both(p, c, prey);
    ORC_TERMS(growth, death, grazed);
    return growth - death - grazed;
}
static double phyto_chl(PP p, PH ph, PC c, double I, double IChl, double IFe, int prey) { /* :59-75, mixed_mondo.jl:112-124 */
    double mu, rho;
    production_and_energy_assimilation_absorption_ratio(p, ph, c, I, IChl, IFe, &mu, &rho);
    double t0 = ph->minimum_chlorophyll_ratio, t1 = ph->maximum_chlorophyll_ratio;
    double growth = (1 - ph->exudated_fraction) * 12 * (t0 + (t1 - t0) * rho) * mu * I;
    double tChl = IChl / (12 * I + EPS0);
    double l, q;
    phyto_mortality(p, ph, c, I, IChl, IFe, &l, &q);
    double death = (l + q);
    double grazed = grazing_both(p, c, prey);
    ORC_TERMS(growth, death * tChl * 12, grazed * tChl * 12);
    return growth - (death + grazed) * tChl * 12;
}
static double phyto_iron(PP p, PH ph, PC c, double I, double IChl, double IFe, int prey) { /* :77-93 */
    double growth = iron_uptake(p, ph, c, I, IChl, IFe);
    double tFe = IFe / (I + EPS0);
    double l, q;
    phyto_mortality(p, ph, c, I, IChl, IFe, &l, &q);
    double death = (l + q);
    double grazed = grazing_both(p, c, prey);
    ORC_TERMS(growth, death * tFe, grazed * tFe);
    return growth - (death + grazed) * tFe;
}

static double tendency(PP p, PC c, int name) {
    switch (name) {
        case T_P: return phyto_carbon(p, NANO, 0);
        case T_D: return phyto_carbon(p, DIAT, 1);
        case T_PChl: return phyto_chl(p, NANO, 0);
        case T_DChl: return phyto_chl(p, DIAT, 1);
        case T_PFe: return phyto_iron(p, NANO, 0);
        case T_DFe: return phyto_iron(p, DIAT, 1);
        case T_DSi: { /* mixed_mondo_nano_diatoms.jl:95-112 */
            double growth = silicate_uptake(p, DIAT);
            double tSi = c->DSi / (c->D + EPS0);
            double l, q;
            phyto_mortality(p, DIAT, &l, &q);
            double death = (l + q);
            double grazed = grazing_both(p, c, 1);
            ORC_TERMS(growth, death * tSi, grazed * tSi);
            return growth - (death + grazed) * tSi;
        }
        case T_Z: { /* micro_and_meso.jl:36-48: M preys on Z */
            double net = growth_death(p, &p->micro, c, 0);
            double s_net = orc_term_scale;
            double predatory = zoo_grazing_on(p, &p->meso, c, 1, 3);
            ORC_TERMS(s_net, predatory);
            return net - predatory;
        }
        case T_M: return growth_death(p, &p->meso, c, 1) - 0.0; /* scale: growth_death's */
        case T_DOC: { /* dissolved_organic_carbon.jl:39-54 */
            double exud = p->nano.exudated_fraction * total_production(p, NANO) + p->diatoms.exudated_fraction * total_production(p, DIAT);
            double ute = upper_trophic_excretion(p, c);
            double gw = organic_excretion(p, c);
            double pb = specific_degradation_rate(p, c) * c->POC;
            double db = dom_degradation(p, c);
            double agg;
            dom_aggregation(p, c, &agg, NULL, NULL, NULL);
            ORC_TERMS(exud, ute, gw, pb, db, agg);
            return (exud + ute + gw + pb - db - agg);
        }
        case T_POC: { /* particulate_organic_matter/carbon.jl:3-26 */
            double gw = non_assimilated_waste(p, &p->micro, c, 0);
            double pm = small_mortality_phyto(p, c);
            double zm = zoo_mortality(p, &p->micro, c, 0);
            double F1, F3;
            dom_aggregation(p, c, NULL, &F1, NULL, &F3);
            double da = F1 + F3;
            double lb = specific_degradation_rate(p, c) * c->GOC;
            double gr = total_grazing_POC(p, c);
            double atl = pom_aggregation(p, c);
            double sb = specific_degradation_rate(p, c) * c->POC;
            ORC_TERMS(gw, pm, zm, da, lb, gr, atl, sb);
            return (gw + pm + zm + da + lb - gr - atl - sb);
        }
        case T_GOC: { /* carbon.jl:28-50 */
            double gw = non_assimilated_waste(p, &p->meso, c, 1);
            double pm = large_mortality_phyto(p, c);
            double zm = zoo_linear_mortality(p, &p->meso, c, 1);
            double atl = pom_aggregation(p, c);
            double utf = upper_trophic_fecal_production(p, c);
            double F2;
            dom_aggregation(p, c, NULL, NULL, &F2, NULL);
            double gr = total_grazing_GOC(p, c);
            double lb = specific_degradation_rate(p, c) * c->GOC;
            ORC_TERMS(gw, pm, zm, utf, atl, F2, gr, lb);
            return (gw + pm + zm + utf + atl + F2 - gr - lb);
        }
        case T_SFe: { /* particulate_organic_matter/iron.jl:2-45 */
            double theta = c->SFe / (c->POC + EPS0);
            double gw = non_assimilated_iron_waste(p, &p->micro, c, 0);
            double pm = small_mortality_iron_phyto(p, c);
            double zm = zoo_mortality(p, &p->micro, c, 0) * p->micro.iron_ratio;
            double lb = specific_degradation_rate(p, c) * c->BFe;
            double lFe = iron_scavenging_rate(p, c);
            double Fep = free_iron(c);
            double scav = lFe * c->POC * Fep;
            double ba = p->small_fraction_of_bacterially_consumed_iron * bacterial_iron_uptake(p, c);
            double ca;
            aggregation_of_colloidal_iron(p, c, NULL, &ca, NULL);
            double gr = total_grazing_POC(p, c) * theta;
            double atl = pom_aggregation(p, c) * theta;
            double sb = specific_degradation_rate(p, c) * c->SFe;
            ORC_TERMS(gw, pm, zm, lb, lFe * c->POC * orc_fep_scale, ba, orc_cg_scale[0], gr, atl, sb);
            return (gw + pm + zm + lb + scav + ba + ca - gr - atl - sb);
        }
        case T_BFe: { /* iron.jl:47-89 */
            double tS = c->SFe / (c->POC + EPS0), tB = c->BFe / (c->GOC + EPS0);
            double gw = non_assimilated_iron_waste(p, &p->meso, c, 1);
            double pm = large_mortality_iron_phyto(p, c);
            double zm = zoo_linear_mortality(p, &p->meso, c, 1) * p->meso.iron_ratio;
            double atl = pom_aggregation(p, c) * tS;
            double utf = upper_trophic_fecal_iron_production(p, c);
            double lFe = iron_scavenging_rate(p, c);
            double Fep = free_iron(c);
            double scav = lFe * c->GOC * Fep;
            double ba = p->large_fraction_of_bacterially_consumed_iron * bacterial_iron_uptake(p, c);
            double ca;
            aggregation_of_colloidal_iron(p, c, NULL, NULL, &ca);
            double gr = total_grazing_GOC(p, c) * tB;
            double lb = specific_degradation_rate(p, c) * c->BFe;
            ORC_TERMS(gw, pm, zm, utf, lFe * c->GOC * orc_fep_scale, ba, orc_cg_scale[1], atl, gr, lb);
            return (gw + pm + zm + utf + scav + ba + ca + atl - gr - lb);
        }
        case T_PSi: { /* particulate_organic_matter/silicate.jl:1-9 */
            double prod = particulate_silicate_production(p, c), diss = particulate_silicate_dissolution(p, c);
            ORC_TERMS(prod, diss);
            return prod - diss;
        }
        case T_CaCO3: { /* particulate_organic_matter/calcite.jl:1-7 */
            double prod = calcite_production(p, c), diss = calcite_dissolution(p, c);
            ORC_TERMS(prod, diss);
            return prod - diss;
        }
        case T_NO3: { /* nitrate_ammonia.jl:22-32 */
            double nitrif = nitrification(p, c);
            double remin = oxic_remineralisation(p, c);
            double consumption = uptake_NO3(p, NANO) + uptake_NO3(p, DIAT);
            ORC_TERMS(nitrif, p->nitrogen_redfield_ratio * remin, p->nitrogen_redfield_ratio * consumption);
            return nitrif + p->nitrogen_redfield_ratio * (remin - consumption);
        }
        case T_NH4: { /* nitrate_ammonia.jl:34-50 */
            double nitrif = nitrification(p, c);
            double remin = anoxic_remineralisation(p, c);
            double consumption = uptake_NH4(p, NANO) + uptake_NH4(p, DIAT);
            double gw = inorganic_excretion(p, c);
            double utw = upper_trophic_respiration(p, c);
            double fix = nitrogen_fixation(p, c);
            ORC_TERMS(fix, p->nitrogen_redfield_ratio * remin, p->nitrogen_redfield_ratio * gw, p->nitrogen_redfield_ratio * utw,
                      p->nitrogen_redfield_ratio * consumption, nitrif);
            return fix + p->nitrogen_redfield_ratio * (remin + gw + utw - consumption) - nitrif;
        }
        case T_PO4: { /* phosphate.jl:21-33 */
            double up = total_production(p, NANO) + total_production(p, DIAT);
            double gw = inorganic_excretion(p, c);
            double rp = upper_trophic_respiration(p, c);
            double remin = dom_degradation(p, c);
            ORC_TERMS(p->phosphate_redfield_ratio * gw, p->phosphate_redfield_ratio * rp, p->phosphate_redfield_ratio * remin,
                      p->phosphate_redfield_ratio * up);
            return p->phosphate_redfield_ratio * (gw + rp + remin - up);
        }
        case T_Fe: { /* iron/simple_iron.jl:19-53 */
            double lFe = iron_scavenging_rate(p, c);
            double Fep = free_iron(c);
            double Lt = p->dissolved_ligand_ratio * c->DOC - p->maximum_ligand_concentration; /* :55-62 */
            double total_ligand = jl_max(p->maximum_ligand_concentration, Lt);
            double ligand_aggregation = p->excess_scavenging_enhancement * lFe * jl_max(0, c->Fe - total_ligand) * Fep;
            double colloidal;
            aggregation_of_colloidal_iron(p, c, &colloidal, NULL, NULL);
            double scav = iron_scavenging(p, c);
            double BactFe = bacterial_iron_uptake(p, c);
            double small_particles = specific_degradation_rate(p, c) * c->SFe;
            double consumption = iron_uptake(p, NANO) + iron_uptake(p, DIAT);
            double gw = non_assimilated_iron(p, c);
            double s_gw = orc_term_scale; /* itself intake − waste − growth: its own Σ|terms| */
            double utw = upper_trophic_dissolved_iron(p, c);
            ORC_TERMS(small_particles, s_gw, utw, consumption,
                      p->excess_scavenging_enhancement * lFe * jl_max(0, c->Fe - total_ligand) * orc_fep_scale,
                      orc_cg_scale[0] + orc_cg_scale[1], lFe * (c->POC + c->GOC) * orc_fep_scale, BactFe);
            return (small_particles + gw + utw - consumption - ligand_aggregation - colloidal - scav - BactFe);
        }
        case T_Si: { /* silicate.jl:20-26 */
            double diss = particulate_silicate_dissolution(p, c), up = silicate_uptake(p, DIAT);
            ORC_TERMS(diss, up);
            return diss - up;
        }
        case T_DIC: { /* inorganic_carbon.jl:32-47 */
            double zr = inorganic_excretion(p, c);
            double ut = upper_trophic_respiration(p, c);
            double remin = dom_degradation(p, c);
            double cd = calcite_dissolution(p, c);
            double cp = calcite_production(p, c);
            double consumption = total_production(p, NANO) + total_production(p, DIAT);
            ORC_TERMS(zr, ut, remin, cd, cp, consumption);
            return (zr + ut + remin + cd - cp - consumption);
        }
        case T_Alk: { /* inorganic_carbon.jl:49-58 */
            double nitrate_production = tendency(p, c, T_NO3);
            double s_no3 = orc_term_scale;
            double ammonia_production = tendency(p, c, T_NH4);
            double s_nh4 = orc_term_scale;
            double calcite_prod = tendency(p, c, T_CaCO3);
            double s_ca = orc_term_scale;
            ORC_TERMS(s_nh4, s_no3, 2 * s_ca); /* a sum of three tendencies: the un-cancelled scale is the sum of theirs */
            return ammonia_production - nitrate_production - 2 * calcite_prod;
        }
        case T_O2: { /* oxygen.jl:30-51 */
            double tr = p->ratio_for_respiration, tn = p->ratio_for_nitrification;
            double zoo = tr * inorganic_excretion(p, c);
            double ut = tr * upper_trophic_respiration(p, c);
            double remin = ((tr + tn) * oxic_remineralisation(p, c) + tr * anoxic_remineralisation(p, c));
            double ap = tr * (uptake_NH4(p, NANO) + uptake_NH4(p, DIAT));
            double np = (tr + tn) * (uptake_NO3(p, NANO) + uptake_NO3(p, DIAT));
            double nitrif = tn * nitrification(p, c) / p->nitrogen_redfield_ratio;
            double fix = tn * nitrogen_fixation(p, c) / p->nitrogen_redfield_ratio;
            ORC_TERMS(ap, np, fix, remin, zoo, ut, nitrif);
            return (ap + np + fix - remin - zoo - ut - nitrif);
        }
        default: orc_term_scale = 0.0; return 0.0; /* T, S: zero(grid) PISCES.jl:120 */
    }
}

static void load_pcell(const obm_grid* g, const double* const* tr, const obm_pisces_fields* a, int i, int j, int k, pcell* c) {
    int64_t idx = cell_index(g, i, j, k), pl = plane_index(g, i, j);
    int64_t sz = ((int64_t)g->Nx + 2 * g->Hx) * ((int64_t)g->Ny + 2 * g->Hy);
    double* v = &c->P;
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++) v[n] = tr[n] ? tr[n][idx] : 0.0;
    c->PAR1 = a->PAR1[idx]; c->PAR2 = a->PAR2[idx]; c->PAR3 = a->PAR3[idx]; c->PAR = a->PAR[idx];
    c->Omega = a->Omega[idx];
    c->wPOC = (a->wPOC[idx] + a->wPOC[idx + sz]) / 2;
    c->wGOC = (a->wGOC[idx] + a->wGOC[idx + sz]) / 2;
    c->zmxl = a->mixed_layer_depth_xy[pl];
    c->zeu = a->euphotic_depth_xy[pl];
    c->kappa = a->mean_mixed_layer_vertical_diffusivity_xy[pl];
    c->mlPAR = a->mean_mixed_layer_light_xy[pl];
    c->z = g->zc[k + g->Hz];
}

/* one pass = one tracer (one compute_Gc! launch of the reference) */
int orc_pisces_tendency(const obm_grid* g, const obm_pisces_params* p, const double* const* tracers,
                        const obm_pisces_fields* aux, int name, double* G, int accumulate) {
    if (name < 0 || name >= OBM_PISCES_NTRACERS) return OBM_ESIZE;
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < g->Nz; k++)
        for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
                pcell c;
                load_pcell(g, tracers, aux, i, j, k, &c);
                double t = tendency(p, &c, name);
                int64_t idx = cell_index(g, i, j, k);
                if (accumulate) G[idx] += t; else G[idx] = t;
            }
    return 0;
}

int orc_pisces_tendencies(const obm_grid* g, const obm_pisces_params* p, const double* const* tracers,
                          const obm_pisces_fields* aux, double* const* G, int accumulate) {
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++) {
        if (!G[n]) continue;
        int rc = orc_pisces_tendency(g, p, tracers, aux, n, G[n], accumulate);
        if (rc) return rc;
    }
    return 0;
}

/* Σ|additive terms| of every tendency (the parity metric's S) on the grid: S[n] parent arrays, NULL skips */
int orc_pisces_tendency_scales(const obm_grid* g, const obm_pisces_params* p, const double* const* tracers,
                               const obm_pisces_fields* aux, double* const* S) {
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < g->Nz; k++)
        for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
                pcell c;
                load_pcell(g, tracers, aux, i, j, k, &c);
                int64_t idx = cell_index(g, i, j, k);
                for (int n = 0; n < OBM_PISCES_NTRACERS; n++) {
                    if (!S[n]) continue;
                    orc_term_scale = 0.0;
                    (void)tendency(p, &c, n);
                    S[n][idx] = orc_term_scale;
                }
            }
    return 0;
}

/* scalar entry for box-model style checks: values[26] tracers, aux scalars → tendencies[26] */
void orc_pisces_point(const obm_pisces_params* p, const double* values, double PAR1, double PAR2, double PAR3, double PAR,
                      double Omega, double wPOC, double wGOC, double zmxl, double zeu, double kappa, double mlPAR, double z,
                      double* out) {
    pcell c;
    double* v = &c.P;
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++) v[n] = values[n];
    c.PAR1 = PAR1; c.PAR2 = PAR2; c.PAR3 = PAR3; c.PAR = PAR; c.Omega = Omega; c.wPOC = wPOC; c.wGOC = wGOC;
    c.zmxl = zmxl; c.zeu = zeu; c.kappa = kappa; c.mlPAR = mlPAR; c.z = z;
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++) out[n] = tendency(p, &c, n);
}
/* the same point, also returning Σ|terms| per tendency */
void orc_pisces_point_terms(const obm_pisces_params* p, const double* values, double PAR1, double PAR2, double PAR3, double PAR,
                            double Omega, double wPOC, double wGOC, double zmxl, double zeu, double kappa, double mlPAR, double z,
                            double* out, double* scale) {
    pcell c;
    double* v = &c.P;
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++) v[n] = values[n];
    c.PAR1 = PAR1; c.PAR2 = PAR2; c.PAR3 = PAR3; c.PAR = PAR; c.Omega = Omega; c.wPOC = wPOC; c.wGOC = wGOC;
    c.zmxl = zmxl; c.zeu = zeu; c.kappa = kappa; c.mlPAR = mlPAR; c.z = z;
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++) {
        orc_term_scale = 0.0;
        out[n] = tendency(p, &c, n);
        scale[n] = orc_term_scale;
    }
}

/* Julia's sind / cosd reduce the argument in DEGREES (rem(x, 360) is exact) and fold it to |angle| <= 45° before the
 * radian conversion; sin(x·π/180) would carry the rounding of x·π/180, 2e-11 at the x ~ 1e7 that the reference's
 * swapped day_length(φ, t) passes as the latitude. */
static double jl_sind(double x) {
    double r = fmod(x, 360.0), a = fabs(r), v;
    const double D2R = 3.14159265358979323846 / 180.0;
    if (a < 45.0) v = sin(a * D2R);
    else if (a <= 135.0) v = cos((90.0 - a) * D2R);
    else if (a < 225.0) v = sin((180.0 - a) * D2R);
    else if (a <= 315.0) v = -cos((270.0 - a) * D2R);
    else v = sin((a - 360.0) * D2R);
    return r < 0 ? -v : v;
}
static double jl_cosd(double x) {
    double a = fabs(fmod(x, 360.0));
    const double D2R = 3.14159265358979323846 / 180.0;
    if (a <= 45.0) return cos(a * D2R);
    if (a < 135.0) return sin((90.0 - a) * D2R);
    if (a <= 225.0) return -cos((180.0 - a) * D2R);
    if (a < 315.0) return sin((a - 270.0) * D2R);
    return cos((360.0 - a) * D2R);
}

/* (day_length::CBMDayLength)(t, φ) — src/Utils/Utils.jl:13-34 */
double orc_cbm_day_length(double t, double phi) {
    const double pcoef = 0.833;
    const double D2R = 3.14159265358979323846 / 180.0;
    double J = floor(fmod(t, 365 * DAY) / DAY); /* mod(t, 365days) for t >= 0 */
    if (fmod(t, 365 * DAY) < 0) J = floor((fmod(t, 365 * DAY) + 365 * DAY) / DAY);
    double theta = 0.216310 + 2 * atan(0.9671396 * tan(0.00860 * (J - 186)));
    double decl = asin(0.39795 * cos(theta)) / D2R; /* asind */
    double L = jl_max(-1.0, jl_min(1.0, (jl_sind(pcoef) + jl_sind(phi) * jl_sind(decl)) / (jl_cosd(phi) * jl_cosd(decl))));
    return (24 - 24.0 / 180 * (acos(L) / D2R)) * 3600.0;
}
