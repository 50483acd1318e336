// pisces_tendencies.cu — ONE fused kernel for the 24 PISCES tracer tendencies.
//
// Replaces the 24 per-tracer callables of src/Models/AdvectedPopulations/PISCES/ which
// Oceananigans evaluates in 26 compute_Gc! launches per stage, each re-reading its
// inputs and re-evaluating nutrient limitation / growth rates / grazing / mortalities / bacteria /
// aggregation / iron chemistry (≈ 24 × 15 field reads and ≥ 10× redundant FP64 work per cell).
// Here a thread owns a cell: 23 tracers (DIC, Alk, S are never read) + PAR₁₂₃, PAR, Ω, w (two
// faces of two fields) are read once, coalesced along x; every shared sub-model is evaluated once
// in registers in the reference's operation order; temperature powers b^T are evaluated once per
// DISTINCT base; each tendency is stored as soon as its last term is known.
//
// FP64-pipe bound: ≈ 10 exp-class calls + 1 pow + 1 sqrt + ≈ 60 divisions per cell (vs ≈ 110
// divisions and ≈ 90 transcendental calls of the un-fused reference path).
// Reference quirks are reproduced on purpose: swapped day-length arguments
// (host-evaluated, obm_pisces_params), (P, D, POC, Z) prey order, negative flux-feeding "flux",
// Alk = NH₄ − NO₃ − 2·CaCO₃.
#include <string.h>

#include "obm_common.cuh"

namespace obm {

enum { T_P = 0, T_PChl, T_PFe, T_D, T_DChl, T_DFe, T_DSi, T_Z, T_M, T_DOC, T_POC, T_GOC, T_SFe, T_BFe, T_PSi, T_CaCO3,
       T_NO3, T_NH4, T_PO4, T_Fe, T_Si, T_DIC, T_Alk, T_O2, T_T, T_S };

struct PiscesArgs {
    GridDims d;
    obm_pisces_params p;
    obm_pisces_fields f;
    const double* c[OBM_PISCES_NTRACERS];
    double* g[OBM_PISCES_NTRACERS];
    // temperature bases, users: 0 nano, 1 diatoms, 2 micro, 3 meso, 4 DOM, 5 POM
    double ln_base[6];
    int same_as[6];  // index of an earlier user with an identical base, or -1
    int accumulate;
};

constexpr double DAY = 86400.0;

__device__ __forceinline__ double min3(double a, double b, double c) { return jl_min(jl_min(a, b), c); }
__device__ __forceinline__ double min4(double a, double b, double c, double d) { return jl_min(jl_min(jl_min(a, b), c), d); }

struct Cell {
    double NO3, NH4, PO4, Fe, Si, T, O2;
    double PAR1, PAR2, PAR3;
    double zmxl, zeu, kappa, z;
};

struct Phyto {  // everything the rest of the model needs from one phytoplankton class
    double L, LFe, LPO4, LN, LNO3, LNH4;
    double mu, muI, lin, quad, tFe, tChl, mui;
};

// nutrient_limitation.jl:20-73 + growth_rate.jl:3-47 + mixed_mondo.jl:137-175, evaluated ONCE per class
__device__ __forceinline__ Phyto phytoplankton(const obm_pisces_params& p, const obm_pisces_phyto& ph, const Cell& c,
                                               double I, double IChl, double IFe, double fT, double shear) {
    Phyto r;
    // quotas
    r.tFe = IFe / (I + eps0());
    r.tChl = IChl / (12 * I + eps0());
    const double tFe_l = I == 0 ? 0.0 : r.tFe;
    const double tChl_l = I == 0 ? 0.0 : r.tChl;
    // size_factor mixed_mondo.jl:207-215
    const double I1 = jl_min(I, ph.threshold_for_size_dependency);
    const double I2 = jl_max(0.0, I - ph.threshold_for_size_dependency);
    const double Kbar = (I1 + ph.size_ratio * I2) / (I1 + I2 + eps0());
    const double Kno = ph.minimum_nitrate_half_saturation * Kbar, Knh = ph.minimum_ammonium_half_saturation * Kbar;
    const double Kp = ph.minimum_phosphate_half_saturation * Kbar, Ksi = ph.minimum_silicate_half_saturation * Kbar;
    // nitrogen_limitation(N₁, N₂, K₁, K₂) nutrient_limitation.jl:73
    r.LNO3 = (Knh * c.NO3) / (Kno * Knh + Kno * c.NH4 + Knh * c.NO3 + eps0());
    r.LNH4 = (Kno * c.NH4) / (Knh * Kno + Knh * c.NO3 + Kno * c.NH4 + eps0());
    r.LN = r.LNO3 + r.LNH4;
    r.LPO4 = c.PO4 / (c.PO4 + Kp + eps0());
    const double tm = 1000 * (0.0016 / 55.85 * 12 * tChl_l + 1.5 * 1.21e-5 * 14 / (55.85 * 7.625) * r.LN
                              + 1.15e-4 * 14 / (55.85 * 7.625) * r.LNO3);
    r.LFe = jl_min(1.0, jl_max(0.0, (tFe_l - tm) / ph.optimal_iron_quota));
    const double Sip = p.silicate_climatology, pk = ph.silicate_half_saturation_parameter;
    const double KSi = Ksi + 7 * (Sip * Sip) / (pk * pk + Sip * Sip);
    double LSi = c.Si / (c.Si + KSi);
    LSi = ph.silicate_limited ? LSi : __longlong_as_double(0x7ff0000000000000LL);
    r.L = min4(r.LN, r.LPO4, r.LFe, LSi);

    // growth rate (μ::BaseProduction)(…, L) with the SWAPPED day length — growth_rate.jl:3-47
    const double PAR = ph.blue_light_absorption * c.PAR1 + ph.green_light_absorption * c.PAR2 + ph.red_light_absorption * c.PAR3;
    const double dl = p.day_length_growth;
    const double dd = jl_max(0.0, c.zeu - c.zmxl);
    const double drt = dd * dd / c.kappa;
    r.mui = ph.base_growth_rate * fT;
    const double f1 = 1.5 * dl / (dl + 0.5 * DAY);
    const double f2 = 1 - drt / (drt + ph.dark_tolerance);
    double alpha = ph.initial_slope_of_PI_curve;
    if (ph.low_light_adaptation != 0.0) alpha = alpha * (1 + ph.low_light_adaptation * exp(-PAR));
    else alpha = alpha * (1 + 0.0);
    double fl;
    if (ph.growth_rate_kind == OBM_GROWTH_NUTRIENT_LIMITED)
        fl = 1 - exp(-alpha * r.tChl * PAR / (dl * r.mui * r.L + eps0()));
    else
        fl = 1 - exp(-alpha * r.tChl * PAR / (dl * (ph.basal_respiration_rate + ph.reference_growth_rate)));
    r.mu = r.mui * f1 * f2 * fl * r.L;
    r.muI = r.mu * I;

    // mortality mixed_mondo.jl:137-167
    r.lin = ph.linear_mortality_rate * I / (I + ph.mortality_half_saturation) * I;
    const double w = ph.base_quadratic_mortality + ph.maximum_quadratic_mortality * 0.25 * (1 - r.L * r.L) / (0.25 + r.L * r.L);
    r.quad = shear * w * (I * I);
    return r;
}

// chlorophyll synthesis: production_and_energy_assimilation_absorption_ratio (growth_rate.jl:126-156)
// + chlorophyll_growth (mixed_mondo.jl:112-124); CORRECT day-length order here
__device__ __forceinline__ double chlorophyll_growth(const obm_pisces_params& p, const obm_pisces_phyto& ph, const Cell& c,
                                                     const Phyto& r, double I, double IChl) {
    const double PAR = ph.blue_light_absorption * c.PAR1 + ph.green_light_absorption * c.PAR2 + ph.red_light_absorption * c.PAR3;
    const double dl = p.day_length_chlorophyll;
    const double f1 = 1.5 * dl / (dl + 0.5 * DAY);
    const double mucheck = r.mu / f1 * dl;
    double alpha = ph.initial_slope_of_PI_curve;
    if (ph.low_light_adaptation != 0.0) alpha = alpha * (1 + ph.low_light_adaptation * exp(-PAR));
    else alpha = alpha * (1 + 0.0);
    const double rho = 12 * mucheck * I / (alpha * IChl * PAR + eps0()) * r.L;
    const double t0 = ph.minimum_chlorophyll_ratio, t1 = ph.maximum_chlorophyll_ratio;
    return (1 - ph.exudated_fraction) * 12 * (t0 + (t1 - t0) * rho) * r.mu * I;
}

// iron_uptake mixed_mondo.jl:177-205
__device__ __forceinline__ double iron_uptake(const obm_pisces_phyto& ph, const Cell& c, const Phyto& r, double I) {
    const double I1 = jl_min(I, ph.threshold_for_size_dependency);
    const double I2 = jl_max(0.0, I - ph.threshold_for_size_dependency);
    const double K = ph.half_saturation_for_iron_uptake * ((I1 + ph.size_ratio * I2) / (I1 + I2 + eps0()));
    const double L1 = c.Fe / (c.Fe + K + eps0());
    const double L2 = 4 - 4.5 * r.LFe / (r.LFe + 1);
    const double q = r.tFe / ph.maximum_iron_ratio;
    return (1 - ph.exudated_fraction) * ph.maximum_iron_ratio * L1 * L2 * jl_max(0.0, (1 - q) / (1.05 - q)) * r.mui * I;
}

struct Zoo {
    double tsg, avail, ge, gI, gfI, base_ff, inv_avail_I;  // inv_avail_I: tsg / (avail + eps) handled per use
    double mort, lin_mort, iron_graze, iron_ff;
};

// food_quality_dependant.jl:126-220, iron_grazing.jl:2-51 — evaluated ONCE per class
template <int N>
__device__ __forceinline__ Zoo zooplankton(const obm_pisces_params& p, const obm_pisces_zoo& z, const double (&food)[4],
                                           const double (&iron)[4], double I, double fT, double dO2, double flux_C,
                                           double flux_Fe) {
    Zoo r;
    const double J = z.specific_food_threshold_concentration;
    const double base = z.maximum_grazing_rate * fT;
    double total_food = food[0] * z.food_preferences[0];
    double avail = jl_max(0.0, (food[0] - J)) * z.food_preferences[0];
    double total_iron = iron[0] * z.food_preferences[0];
    double s = jl_max(0.0, (food[0] - J)) * z.food_preferences[0] * iron[0];
#pragma unroll
    for (int n = 1; n < N; n++) {
        total_food += food[n] * z.food_preferences[n];
        const double a = jl_max(0.0, (food[n] - J)) * z.food_preferences[n];
        avail += a;
        total_iron += iron[n] * z.food_preferences[n];
        s += a * iron[n];
    }
    const double clg = jl_max(0.0, avail - jl_min(avail / 2, z.food_threshold_concentration));
    r.tsg = base * clg / (z.grazing_half_saturation + total_food);
    r.avail = avail;
    const double igr = total_iron / (z.iron_ratio * r.tsg + eps0());
    r.ge = jl_min(1.0, igr) * jl_min(z.minimum_growth_efficiency, (1 - z.non_assimilated_fraction) * igr);
    r.gI = r.tsg * I;
    r.base_ff = z.maximum_flux_feeding_rate * fT;
    r.gfI = r.base_ff * flux_C * I;
    const double cf = I / (I + z.mortality_half_saturation);
    r.mort = fT * I * (z.quadratic_mortality * I + z.linear_mortality * (cf + 3 * dO2));
    r.lin_mort = fT * z.linear_mortality * (cf + 3 * dO2) * I;
    r.iron_graze = s * r.tsg / (avail + eps0()) * I;
    r.iron_ff = r.base_ff * flux_Fe * I;
    return r;
}
// grazing on one prey — food_quality_dependant.jl:226-255
__device__ __forceinline__ double graze_on(const obm_pisces_zoo& z, const Zoo& r, double pref, double prey, double I) {
    return pref * jl_max(0.0, prey - z.specific_food_threshold_concentration) * r.tsg / (r.avail + eps0()) * I;
}

__device__ __forceinline__ void put(double* g, long long idx, double t, int accumulate) {
    if (g == nullptr) return;
    if (accumulate) t += g[idx];
    g[idx] = t;
}

__global__ void __launch_bounds__(128) pisces_tendency_kernel(const __grid_constant__ PiscesArgs a) {
    int i, j, k;
    if (!thread_cell(a.d, i, j, k)) return;
    const long long idx = cell_index(a.d, i, j, k);
    const long long pl = plane_index(a.d, i, j);
    const obm_pisces_params& p = a.p;
    const int acc = a.accumulate;

    // ---- one coalesced read of the cell ------------------------------------------------------------
    const double P = a.c[T_P][idx], PChl = a.c[T_PChl][idx], PFe = a.c[T_PFe][idx];
    const double D = a.c[T_D][idx], DChl = a.c[T_DChl][idx], DFe = a.c[T_DFe][idx], DSi = a.c[T_DSi][idx];
    const double Z = a.c[T_Z][idx], M = a.c[T_M][idx], DOC = a.c[T_DOC][idx];
    const double POC = a.c[T_POC][idx], GOC = a.c[T_GOC][idx], SFe = a.c[T_SFe][idx], BFe = a.c[T_BFe][idx];
    const double PSi = a.c[T_PSi][idx], CaCO3 = a.c[T_CaCO3][idx];
    Cell c;
    c.NO3 = a.c[T_NO3][idx]; c.NH4 = a.c[T_NH4][idx]; c.PO4 = a.c[T_PO4][idx]; c.Fe = a.c[T_Fe][idx];
    c.Si = a.c[T_Si][idx]; c.O2 = a.c[T_O2][idx]; c.T = a.c[T_T][idx];
    c.PAR1 = a.f.PAR1[idx]; c.PAR2 = a.f.PAR2[idx]; c.PAR3 = a.f.PAR3[idx];
    const double PARt = a.f.PAR[idx], Omega = a.f.Omega[idx];
    // ℑzᵃᵃᶜ(i, j, k, grid, w) = (w[k] + w[k+1]) / 2 — two_size_class.jl:95-98
    const double wPOC = (a.f.wPOC[idx] + a.f.wPOC[idx + a.d.sz]) / 2;
    const double wGOC = (a.f.wGOC[idx] + a.f.wGOC[idx + a.d.sz]) / 2;
    c.zmxl = a.f.mixed_layer_depth_xy[pl];
    c.zeu = a.f.euphotic_depth_xy[pl];
    c.kappa = a.f.mean_mixed_layer_vertical_diffusivity_xy[pl];
    const double mlPAR = a.f.mean_mixed_layer_light_xy[pl];
    c.z = a.d.zc[k];

    // ---- shared scalars --------------------------------------------------------------------------------
    const double shear = c.z < c.zmxl ? p.background_shear : p.mixed_layer_shear;
    const double dO2 = jl_min(1.0, jl_max(0.0, 0.4 * (p.first_anoxia_threshold - c.O2) / (p.second_anoxia_threshold + c.O2)));
    // b^T once per distinct base (exp(T ln b); bases are parameters, ln b is host-evaluated)
    double fT[6];
#pragma unroll
    for (int u = 0; u < 6; u++) {
        double v = 0.0;
        bool found = false;
#pragma unroll
        for (int q = 0; q < u; q++)
            if (a.same_as[u] == q) { v = fT[q]; found = true; }
        fT[u] = found ? v : exp(c.T * a.ln_base[u]);
    }

    // ---- phytoplankton ------------------------------------------------------------------------------------
    const Phyto n = phytoplankton(p, p.nano, c, P, PChl, PFe, fT[0], shear);
    const Phyto d = phytoplankton(p, p.diatoms, c, D, DChl, DFe, fT[1], shear);

    // ---- zooplankton ----------------------------------------------------------------------------------------
    const double fluxPOC = POC * wPOC, fluxGOC = GOC * wGOC, fluxSFe = SFe * wPOC, fluxBFe = BFe * wGOC;
    const double tSFe = SFe / (POC + eps0());
    const double food[4] = {P, D, POC, Z};
    const double iron[4] = {n.tFe, d.tFe, tSFe, p.micro.iron_ratio};
    const Zoo zz = zooplankton<3>(p, p.micro, food, iron, Z, fT[2], dO2, fluxPOC + fluxGOC, fluxSFe + fluxBFe);
    const Zoo zm = zooplankton<4>(p, p.meso, food, iron, M, fT[3], dO2, fluxPOC + fluxGOC, fluxSFe + fluxBFe);
    // grazing(zoo::MicroAndMeso, prey) = micro + meso (micro_and_meso.jl:50-52)
    const double gP_micro = graze_on(p.micro, zz, p.micro.food_preferences[0], P, Z);
    const double gP_meso = graze_on(p.meso, zm, p.meso.food_preferences[0], P, M);
    const double gP = gP_micro + gP_meso;
    const double gD = graze_on(p.micro, zz, p.micro.food_preferences[1], D, Z) + graze_on(p.meso, zm, p.meso.food_preferences[1], D, M);
    const double gPOC = graze_on(p.micro, zz, p.micro.food_preferences[2], POC, Z) + graze_on(p.meso, zm, p.meso.food_preferences[2], POC, M);
    const double gZ_meso = graze_on(p.meso, zm, p.meso.food_preferences[3], Z, M);

    // ---- P, D, chlorophyll, iron, silicon quotas: mixed_mondo_nano_diatoms.jl:45-112 -----------------
    const double deathP = (n.lin + n.quad), deathD = (d.lin + d.quad);
    put(a.g[T_P], idx, (1 - p.nano.exudated_fraction) * n.muI - deathP - gP, acc);
    put(a.g[T_D], idx, (1 - p.diatoms.exudated_fraction) * d.muI - deathD - gD, acc);
    put(a.g[T_PChl], idx, chlorophyll_growth(p, p.nano, c, n, P, PChl) - (deathP + gP) * n.tChl * 12, acc);
    put(a.g[T_DChl], idx, chlorophyll_growth(p, p.diatoms, c, d, D, DChl) - (deathD + gD) * d.tChl * 12, acc);
    const double upFe_n = iron_uptake(p.nano, c, n, P), upFe_d = iron_uptake(p.diatoms, c, d, D);
    put(a.g[T_PFe], idx, upFe_n - (deathP + gP) * n.tFe, acc);
    put(a.g[T_DFe], idx, upFe_d - (deathD + gD) * d.tFe, acc);
    // silicate_uptake (diatoms) mixed_mondo.jl:217-248
    double upSi;
    {
        const obm_pisces_phyto& ph = p.diatoms;
        const double Si = c.Si, K2 = ph.enhanced_silicate_half_saturation;
        const double L1 = Si / (Si + ph.silicate_half_saturation + eps0());
        const double L2 = p.latitude < 0 ? (Si * Si * Si) / (Si * Si * Si + K2 * K2 * K2) : 0.0;
        const double F1 = min4(d.mu / (d.mui * d.L + eps0()), d.LFe, d.LPO4, d.LN);
        const double F2 = jl_min(1.0, 2.2 * jl_max(0.0, L1 - 0.5));
        const double t1 = ph.optimal_silicate_ratio * L1 * jl_min(5.4, (4.4 * exp(-4.23 * F1) * F2 + 1) * (1 + 2 * L2));
        upSi = (1 - ph.exudated_fraction) * t1 * d.mu * D;
    }
    const double tSi = DSi / (D + eps0());
    put(a.g[T_DSi], idx, upSi - (deathD + gD) * tSi, acc);

    // ---- Z, M: micro_and_meso.jl:36-48 -----------------------------------------------------------------------
    put(a.g[T_Z], idx, (zz.ge * (zz.gI + zz.gfI) - zz.mort) - gZ_meso, acc);
    put(a.g[T_M], idx, (zm.ge * (zm.gI + zm.gfI) - zm.mort) - 0.0, acc);

    // ---- zooplankton wastes: grazing_waste.jl, mortality_waste.jl -------------------------------------------
    const double exc_z = (1 - p.micro.non_assimilated_fraction - zz.ge) * (zz.gI + zz.gfI);
    const double exc_m = (1 - p.meso.non_assimilated_fraction - zm.ge) * (zm.gI + zm.gfI);
    const double inorg_exc = p.micro.dissolved_excretion_fraction * exc_z + p.meso.dissolved_excretion_fraction * exc_m;
    const double org_exc = (1 - p.micro.dissolved_excretion_fraction) * exc_z + (1 - p.meso.dissolved_excretion_fraction) * exc_m;
    const double ut_waste = 1 / (1 - p.meso.minimum_growth_efficiency) * p.meso.quadratic_mortality * fT[3] * (M * M);
    const double ut_R = (1 - p.meso.minimum_growth_efficiency - p.meso.non_assimilated_fraction) * ut_waste;
    const double ut_excretion = (1 - p.meso.dissolved_excretion_fraction) * ut_R;
    const double ut_respiration = p.meso.dissolved_excretion_fraction * ut_R;
    const double ut_fecal = p.meso.non_assimilated_fraction * ut_waste;

    // ---- bacteria: micro_and_meso.jl:85-132 ------------------------------------------------------------------
    const double zmin = jl_min(c.zmxl, c.zeu);
    double Bact;
    {
        const double surface = jl_min(4.0, p.microzooplankton_bacteria_concentration * Z + p.mesozooplankton_bacteria_concentration * M);
        // ifelse(z >= zₘ, 1, (zₘ / z)^a): the discarded arm has no side effect, so it is only evaluated when selected
        const double factor = c.z >= zmin ? 1.0 : pow(zmin / c.z, p.bacteria_concentration_depth_exponent);
        Bact = factor * surface;
    }
    double LBact;
    {
        const double K_NO3 = p.nitrate_half_saturation_for_bacterial_activity, K_NH4 = p.ammonia_half_saturation_for_bacterial_activity;
        const double DOC_limit = DOC / (DOC + p.doc_half_saturation_for_bacterial_activity);
        const double L_N = (K_NO3 * c.NH4 + K_NH4 * c.NO3) / (K_NO3 * K_NH4 + K_NO3 * c.NH4 + K_NH4 * c.NO3);
        const double L_PO4 = c.PO4 / (c.PO4 + p.phosphate_half_saturation_for_bacterial_activity);
        const double L_Fe = c.Fe / (c.Fe + p.iron_half_saturation_for_bacterial_activity);
        LBact = min3(L_N, L_PO4, L_Fe) * DOC_limit;
    }

    // ---- dissolved organic matter: dissolved_organic_carbon.jl:39-130 ---------------------------------------
    const double dom_deg = p.dom_remineralisation_rate * fT[4] * LBact * Bact / p.dom_reference_bacteria_concentration * DOC;
    const double Phi1 = shear * (p.dom_aggregation_parameters[0] * DOC + p.dom_aggregation_parameters[1] * POC) * DOC;
    const double Phi2 = shear * (p.dom_aggregation_parameters[2] * GOC) * DOC;
    const double Phi3 = (p.dom_aggregation_parameters[3] * POC + p.dom_aggregation_parameters[4] * DOC) * DOC;
    const double spec_deg = p.pom_base_breakdown_rate * fT[5] * (1 - 0.45 * dO2);  // two_size_class.jl:127-137
    put(a.g[T_DOC], idx,
        ((p.nano.exudated_fraction * n.muI + p.diatoms.exudated_fraction * d.muI) + ut_excretion + org_exc + spec_deg * POC
         - dom_deg - (Phi1 + Phi2 + Phi3)), acc);

    // ---- iron chemistry: iron/iron.jl:25-37, particulate_organic_matter/iron.jl:97-126 ------------------------
    double Fep;
    {
        const double ligands = jl_max(0.6, 0.09 * (DOC + 40) - 3);
        const double K = exp(16.27 - 1565.7 / jl_max(c.T + 273.15, 5.0));
        const double Dl = 1 + K * ligands - K * c.Fe;
        Fep = (-Dl + sqrt(Dl * Dl + 4 * K * c.Fe)) / (2 * K);
    }
    const double lFe = p.minimum_iron_scavenging_rate + p.load_specific_iron_scavenging_rate * (POC + GOC + CaCO3 + PSi);
    const double BactFe = p.maximum_bacterial_growth_rate * fT[5] * LBact * p.maximum_iron_ratio_in_bacteria * c.Fe
                          / (c.Fe + p.iron_half_saturation_for_bacteria) * Bact * p.bacterial_iron_uptake_efficiency;
    const double colloidal = 0.5 * (c.Fe - Fep);
    const double CgFe1 = (Phi1 + Phi3) * colloidal / (DOC + eps0());
    const double CgFe2 = Phi2 * colloidal / (DOC + eps0());

    // ---- rain ratio & calcite: nano_diatom_coupling.jl:57-124, calcite.jl:9-19 -------------------------------
    double R;
    {
        const double L_CaCO3 = min3(n.LN, c.Fe / (c.Fe + 0.05), n.LPO4);
        const double pcf = jl_max(1.0, P / 2);
        const double low_light = jl_max(0.0, PARt - 1) / (4 + PARt);
        const double high_light = 30 / (30 + PARt);
        const double low_T = jl_max(0.0, c.T / (c.T + 0.1));
        const double high_T = 1 + exp(-((c.T - 10) * (c.T - 10)) / 25);
        const double depth = jl_min(1.0, -50 / c.zmxl);
        R = (p.base_rain_ratio * L_CaCO3 * pcf * low_light * high_light * low_T * high_T * depth);
    }
    const double calcite_loss = p.micro.undissolved_calcite_fraction * gP_micro + p.meso.undissolved_calcite_fraction * gP_meso;
    const double calcite_prod = R * (calcite_loss + (n.lin + n.quad) / 2);
    double calcite_diss;
    {
        const double dCa = jl_max(0.0, 1 - Omega);
        const double e = p.calcite_dissolution_exponent;
        calcite_diss = p.base_calcite_dissolution_rate * (e == 1.0 ? dCa : pow(dCa, e)) * CaCO3;  // x^1.0 ≡ x
    }
    const double tCaCO3 = calcite_prod - calcite_diss;
    put(a.g[T_CaCO3], idx, tCaCO3, acc);

    // ---- POC, GOC: particulate_organic_matter/carbon.jl:3-50 --------------------------------------------------
    const double* ap = p.pom_aggregation_parameters;
    const double pom_agg = shear * (ap[0] * (POC * POC) + ap[1] * POC * GOC) + ap[2] * POC * GOC + ap[3] * (POC * POC);
    const double ff_POC = zz.base_ff * fluxPOC * Z + zm.base_ff * fluxPOC * M;  // flux_feeding(zoo, Val(:POC))
    const double ff_GOC = zz.base_ff * fluxGOC * Z + zm.base_ff * fluxGOC * M;
    const double tg_POC = gPOC + ff_POC;                                      // micro_meso_zoo_coupling.jl:27-32
    const double sm_phyto = (1 - R / 2) * (n.lin + n.quad) + d.lin / 2;         // nano_diatom_coupling.jl:1-9
    const double lm_phyto = R / 2 * (n.lin + n.quad) + d.lin / 2 + d.quad;      // :11-19
    put(a.g[T_POC], idx,
        (p.micro.non_assimilated_fraction * (zz.gI + zz.gfI) + sm_phyto + zz.mort + (Phi1 + Phi3) + spec_deg * GOC
         - tg_POC - pom_agg - spec_deg * POC), acc);
    put(a.g[T_GOC], idx,
        (p.meso.non_assimilated_fraction * (zm.gI + zm.gfI) + lm_phyto + zm.lin_mort + ut_fecal + pom_agg + Phi2
         - ff_GOC - spec_deg * GOC), acc);

    // ---- SFe, BFe: particulate_organic_matter/iron.jl:2-89 ------------------------------------------------------
    {
        const double smi = (1 - R / 2) * (n.lin + n.quad) * n.tFe + d.lin * d.tFe / 2;           // nano_diatom_coupling.jl:21-37
        const double lmi = R / 2 * (n.lin + n.quad) * n.tFe + (d.lin / 2 + d.quad) * d.tFe;       // :39-55
        const double tB = BFe / (GOC + eps0());
        put(a.g[T_SFe], idx,
            (p.micro.non_assimilated_fraction * (zz.iron_graze + zz.iron_ff) + smi + zz.mort * p.micro.iron_ratio + spec_deg * BFe
             + lFe * POC * Fep + p.small_fraction_of_bacterially_consumed_iron * BactFe + CgFe1
             - tg_POC * tSFe - pom_agg * tSFe - spec_deg * SFe), acc);
        put(a.g[T_BFe], idx,
            (p.meso.non_assimilated_fraction * (zm.iron_graze + zm.iron_ff) + lmi + zm.lin_mort * p.meso.iron_ratio
             + ut_fecal * p.meso.iron_ratio + lFe * GOC * Fep + p.large_fraction_of_bacterially_consumed_iron * BactFe + CgFe2
             + pom_agg * tSFe - ff_GOC * tB - spec_deg * BFe), acc);
    }

    // ---- PSi, Si: particulate_organic_matter/silicate.jl:1-48, silicate.jl:20-26 -------------------------------
    double psi_diss;
    {
        const double ll = p.fast_dissolution_rate_of_silicate, lr = p.slow_dissolution_rate_of_silicate;
        const double chi = p.base_liable_silicate_fraction * (c.z >= zmin ? 1.0 : exp((ll - lr) * (zmin - c.z) / wGOC));
        const double l0 = chi * ll + (1 - chi) * lr;
        const double eq = exp10(6.44 - 968 / (c.T + 273.15));
        const double sat = (eq - c.Si) / eq;
        const double q = 1 + c.T / 400;
        const double q2 = q * q;
        const double b = (q2 * q2) * sat;  // ((1 + T/400)^4 * saturation)
        const double b2 = b * b, b4 = b2 * b2;
        const double l = l0 * (0.225 * (1 + c.T / 15) * sat + 0.775 * (b4 * b4 * b));  // (…)^9
        psi_diss = l * PSi;
    }
    put(a.g[T_PSi], idx, (gD + d.lin + d.quad) * tSi - psi_diss, acc);
    put(a.g[T_Si], idx, psi_diss - upSi, acc);

    // ---- nitrogen: nitrogen/nitrate_ammonia.jl:22-89 ------------------------------------------------------------
    const double nitrif = p.maximum_nitrification_rate * c.NH4 / (1 + mlPAR) * (1 - dO2);
    double fixation;
    {
        const double limit = n.LN >= 0.8 ? 0.01 : 1 - n.LN;
        const double growth_requirement = jl_max(0.0, n.mui - 2.15);
        const double nutrient = jl_min(c.Fe / (c.Fe + p.iron_half_saturation_for_fixation),
                                       c.PO4 / (c.PO4 + p.phosphate_half_saturation_for_fixation));
        const double light = 1 - exp(-PARt / p.light_saturation_for_fixation);
        fixation = p.maximum_fixation_rate * growth_requirement * limit * nutrient * light;
    }
    const double upNO3 = n.muI * n.LNO3 / (n.LN + eps0()) + d.muI * d.LNO3 / (d.LN + eps0());
    const double upNH4 = n.muI * n.LNH4 / (n.LN + eps0()) + d.muI * d.LNH4 / (d.LN + eps0());
    const double oxic = (1 - dO2) * dom_deg, anoxic = dO2 * dom_deg;
    const double tN = p.nitrogen_redfield_ratio;
    const double tNO3 = nitrif + tN * (oxic - upNO3);
    const double tNH4 = fixation + tN * (anoxic + inorg_exc + ut_respiration - upNH4) - nitrif;
    put(a.g[T_NO3], idx, tNO3, acc);
    put(a.g[T_NH4], idx, tNH4, acc);

    // ---- PO₄, Fe, DIC, Alk, O₂ ---------------------------------------------------------------------------------------
    const double prod = n.muI + d.muI;
    put(a.g[T_PO4], idx, p.phosphate_redfield_ratio * (inorg_exc + ut_respiration + dom_deg - prod), acc);  // phosphate.jl:21-33
    {   // iron/simple_iron.jl:19-62
        const double Lt = p.dissolved_ligand_ratio * DOC - p.maximum_ligand_concentration;
        const double ligand_agg = p.excess_scavenging_enhancement * lFe * jl_max(0.0, c.Fe - jl_max(p.maximum_ligand_concentration, Lt)) * Fep;
        // non_assimilated_iron grazing_waste.jl:45-63, per class
        const double fz = zz.iron_graze + zz.iron_ff, fm = zm.iron_graze + zm.iron_ff;
        const double nai = (fz - p.micro.non_assimilated_fraction * fz - p.micro.iron_ratio * zz.ge * (zz.gI + zz.gfI))
                           + (fm - p.meso.non_assimilated_fraction * fm - p.meso.iron_ratio * zm.ge * (zm.gI + zm.gfI));
        put(a.g[T_Fe], idx,
            (spec_deg * SFe + nai + p.meso.iron_ratio * ut_R - (upFe_n + upFe_d) - ligand_agg - (CgFe1 + CgFe2)
             - lFe * (POC + GOC) * Fep - BactFe), acc);
    }
    put(a.g[T_DIC], idx, (inorg_exc + ut_respiration + dom_deg + calcite_diss - calcite_prod - prod), acc);  // inorganic_carbon.jl:32-47
    put(a.g[T_Alk], idx, tNH4 - tNO3 - 2 * tCaCO3, acc);                                                   // :49-58
    {   // oxygen.jl:30-51
        const double tr = p.ratio_for_respiration, tn = p.ratio_for_nitrification;
        const double remin = ((tr + tn) * oxic + tr * anoxic);
        put(a.g[T_O2], idx,
            (tr * upNH4 + (tr + tn) * upNO3 + tn * fixation / tN - remin - tr * inorg_exc - tr * ut_respiration
             - tn * nitrif / tN), acc);
    }
}

}  // namespace obm

using namespace obm;

extern "C" int obm_pisces_tendencies(const obm_grid* grid, const obm_pisces_params* p, const double* const* tracers,
                                     const obm_pisces_fields* aux, double* const* G, int accumulate, void* stream) {
    OBM_REQUIRE(p && tracers && aux && G, OBM_ENULL, "obm_pisces_tendencies: params / tracers / aux / G is NULL");
    static thread_local PiscesArgs tl;  // ~2 KB of kernel arguments, kept off the caller's stack
    PiscesArgs& A = tl;
    memset(&A, 0, sizeof(A));
    int rc = make_dims(grid, &A.d, true);
    if (rc) return rc;
    A.p = *p;
    A.f = *aux;
    OBM_REQUIRE(aux->PAR1 && aux->PAR2 && aux->PAR3 && aux->PAR && aux->Omega && aux->wPOC && aux->wGOC
                    && aux->mixed_layer_depth_xy && aux->euphotic_depth_xy && aux->mean_mixed_layer_vertical_diffusivity_xy
                    && aux->mean_mixed_layer_light_xy,
                OBM_ENULL, "obm_pisces_tendencies: an auxiliary field pointer is NULL");
    OBM_REQUIRE(p->nano.growth_rate_kind >= 0 && p->nano.growth_rate_kind <= 1 && p->diatoms.growth_rate_kind >= 0
                    && p->diatoms.growth_rate_kind <= 1,
                OBM_EENUM, "obm_pisces_tendencies: unknown growth_rate_kind");
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++) {
        const bool read = !(n == T_DIC || n == T_Alk || n == T_S);  // never read by any tendency
        OBM_REQUIRE(!read || tracers[n], OBM_ENULL, "obm_pisces_tendencies: tracers[%d] is NULL", n);
        A.c[n] = tracers[n];
        A.g[n] = (n == T_T || n == T_S) ? nullptr : G[n];  // zero(grid) PISCES.jl:120
    }
    const double bases[6] = {p->nano.temperature_sensitivity, p->diatoms.temperature_sensitivity,
                             p->micro.temperature_sensitivity, p->meso.temperature_sensitivity,
                             p->dom_temperature_sensitivity,   p->pom_temperature_sensitivity};
    for (int u = 0; u < 6; u++) {
        A.ln_base[u] = log(bases[u]);
        A.same_as[u] = -1;
        for (int q = 0; q < u; q++)
            if (bases[q] == bases[u]) { A.same_as[u] = q; break; }
    }
    A.accumulate = accumulate ? 1 : 0;
    const long long cells = cell_count(A.d);
    pisces_tendency_kernel<<<(unsigned)((cells + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A);
    return launch_status("pisces_tendency_kernel");
}
