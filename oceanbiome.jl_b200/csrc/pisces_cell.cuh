// pisces_cell.cuh — the arithmetic of the fused PISCES tendency kernel: one cell in, 24 tendencies out.
//
// Everything here is plain FP64 arithmetic on values already in registers (no loads, no stores, no thread indices), so the
// same source serves three callers: the CUDA kernel (pisces_tendencies.cu — fast pass and exact pass), and, compiled for the
// host by g++ with a few shims (bench_ref/fused_host.cpp), bench.py's "fused-algorithm CPU" baseline and the CPU parity test
// of this very code against the oracle (tests/test_fused_host.py) — the kernel's operation order can be checked without a GPU.
// Reference files are cited per function.
#pragma once

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
    unsigned out_mask;  // bit n set ⇔ g[n] != NULL
    // parameter-only sub-expressions, evaluated once on the host in double precision
    // (the reference re-evaluates them per cell per tracer; results agree to ≤ 1 ulp)
    struct Derived {
        double f1_growth;         // 1.5 dl / (dl + 0.5day), dl = day_length_growth       growth_rate.jl:38
        double dl_over_f1_chl;    // day_length_chlorophyll / f1(day_length_chlorophyll)  growth_rate.jl:145,151
        double inv_resp[2];       // 1 / (dl (bᵣ + μᵣ))                                   growth_rate.jl:123
        double KSi_add[2];        // 7 Si′² / (pk² + Si′²)                                nutrient_limitation.jl:65
        double inv_theta_o[2];    // 1 / optimal_iron_quota
        double inv_theta_Fem[2];  // 1 / maximum_iron_ratio
        double inv_bact_ref;      // 1 / reference_bacteria_concentration
        double ut_coeff;          // 1 / (1 − e₀) · m₀ (meso)                             mortality_waste.jl:40
        double inv_E;             // 1 / light_saturation_for_fixation
        double inv_tN;            // 1 / nitrogen_redfield_ratio
        double K2_cubed;          // enhanced_silicate_half_saturation³
    } dv;
};

constexpr double DAY = 86400.0;
#ifndef OBM_PISCES_BLOCK
#define OBM_PISCES_BLOCK 128
#endif
constexpr int PB = OBM_PISCES_BLOCK;   // threads per block
constexpr int NOUT = 24;  // tendencies

// ---- arithmetic policy -------------------------------------------------------------------------------
// EXACT: IEEE division, NaN-propagating min/max — the reference's semantics operation by operation.
// FAST (default path): a lean branch-free division (MUFU.RCP64H seed + 2 Newton steps, ≤ 1.5 ulp:
// 5 FP64 instructions instead of the ≈ 30-instruction IEEE sequence with its slow-path call
// scaffolding), `x / (y + eps(0.0))` with the reference's exact y == 0 behaviour
// (x·2¹⁰⁷⁴: 0 → 0, finite → ±Inf or the scaled value, NaN → NaN) done by a select, and min/max as
// compare + select.  A cell whose FAST results contain a non-finite value (or whose NaN could be swallowed
// by a min/max) is recomputed with EXACT, so NaN/Inf patterns match the reference everywhere.
#ifndef OBM_PISCES_ROLL
#define OBM_PISCES_ROLL 0  // 1: class-symmetric sub-models as real loops (one copy of their code), see cell_tendencies
#endif
#ifndef OBM_PISCES_EXP
#define OBM_PISCES_EXP 0  // 0: library exp; 1: exp_lean of obm_common.cuh (measured: no gain here, r02); 2: exp_horner
#endif
template <bool EXACT>
struct Ar {
    static constexpr bool EX = EXACT;
    static __device__ __forceinline__ double div(double a, double b) {
        if (EXACT) return a / b;
        return a * rcp_fast(b);  // ≤ 1.5 ulp
    }
    // a / (y + eps(0.0)).  y == 0 makes the quotient garbage (NaN), which the select discards.
    static __device__ __forceinline__ double gdiv(double a, double y) {
        if (EXACT) return a / (y + eps0());
        const double q = div(a, y);
        const double q0 = (a * 0x1p537) * 0x1p537;  // a / 2⁻¹⁰⁷⁴, exact incl. overflow to ±Inf
        return y == 0.0 ? q0 : q;
    }
    // FAST min/max are a compare + select (3 instructions; fmin/fmax on doubles expand to ≈ 7 with their NaN fix-up).
    // NaN: a NaN in `b` propagates, a NaN in `a` is swallowed — never more than fmin/fmax swallow, so the guard of
    // `needs_exact` (inputs that reach the tendencies only through min/max) still covers every such cell.
    static __device__ __forceinline__ double mx(double a, double b) { return EXACT ? jl_max(a, b) : (a > b ? a : b); }
    static __device__ __forceinline__ double mn(double a, double b) { return EXACT ? jl_min(a, b) : (a < b ? a : b); }
    static __device__ __forceinline__ double mn3(double a, double b, double c) { return mn(mn(a, b), c); }
    static __device__ __forceinline__ double mn4(double a, double b, double c, double d) { return mn(mn(mn(a, b), c), d); }
    static __device__ __forceinline__ double ex(double x) {
#if OBM_PISCES_EXP == 0
        return exp(x);
#elif OBM_PISCES_EXP == 1
        return EXACT ? exp(x) : exp_lean(x);
#elif OBM_PISCES_EXP == 2
        return EXACT ? exp(x) : exp_horner(x);
#else
        // 3: the 64-entry table form of obm_common.cuh, branch-free.  Below −707 the argument is clamped (9e-308 for what is
        // a subnormal or 0: every use adds the result to, or multiplies it with, ordinary numbers); above 709 the result is
        // +Inf like the reference's — non-finite, so the cell is redone by the exact pass; NaN → NaN.
        if (EXACT) return exp(x);
        const double v = exp_table(x < -707.0 ? -707.0 : (x > 709.0 ? 709.0 : x));
        return x > 709.0 ? __longlong_as_double(0x7ff0000000000000LL) : v;
#endif
    }
};

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
template <class A>
__device__ __forceinline__ Phyto phytoplankton(const PiscesArgs& a, const int cls, const Cell& c,
                                               double I, double IChl, double IFe, double fT, double shear) {
    const obm_pisces_phyto& ph = cls == 0 ? a.p.nano : a.p.diatoms;
    Phyto r;
    // quotas
    r.tFe = A::gdiv(IFe, I);
    r.tChl = A::gdiv(IChl, 12 * I);
    const double tFe_l = I == 0 ? 0.0 : r.tFe;
    const double tChl_l = I == 0 ? 0.0 : r.tChl;
    // size_factor mixed_mondo.jl:207-215
    const double I1 = A::mn(I, ph.threshold_for_size_dependency);
    const double I2 = A::mx(0.0, I - ph.threshold_for_size_dependency);
    const double Kbar = A::gdiv(I1 + ph.size_ratio * I2, I1 + I2);
    const double Kno = ph.minimum_nitrate_half_saturation * Kbar, Knh = ph.minimum_ammonium_half_saturation * Kbar;
    const double Kp = ph.minimum_phosphate_half_saturation * Kbar, Ksi = ph.minimum_silicate_half_saturation * Kbar;
    // nitrogen_limitation(N₁, N₂, K₁, K₂) nutrient_limitation.jl:73
    r.LNO3 = A::gdiv(Knh * c.NO3, Kno * Knh + Kno * c.NH4 + Knh * c.NO3);
    r.LNH4 = A::gdiv(Kno * c.NH4, Knh * Kno + Knh * c.NO3 + Kno * c.NH4);
    r.LN = r.LNO3 + r.LNH4;
    r.LPO4 = A::gdiv(c.PO4, c.PO4 + Kp);
    const double tm = 1000 * (0.0016 / 55.85 * 12 * tChl_l + 1.5 * 1.21e-5 * 14 / (55.85 * 7.625) * r.LN
                              + 1.15e-4 * 14 / (55.85 * 7.625) * r.LNO3);
    r.LFe = A::mn(1.0, A::mx(0.0, A::EX ? (tFe_l - tm) / ph.optimal_iron_quota : (tFe_l - tm) * a.dv.inv_theta_o[cls]));
    const double KSi = Ksi + a.dv.KSi_add[cls];
    double LSi = A::div(c.Si, c.Si + KSi);
    LSi = ph.silicate_limited ? LSi : __longlong_as_double(0x7ff0000000000000LL);
    // min(L_N, L_PO₄, L_Fe, L_Si) nutrient_limitation.jl:69.  FAST: a select propagates a NaN in its SECOND operand only, and
    // L_Fe is the one limitation a finite state can turn into NaN (a positive biomass far below its pigment makes both
    // quotas overflow: Inf − Inf) — it goes last, so that such a cell reaches the exact pass through a non-finite result.
    r.L = A::EX ? A::mn4(r.LN, r.LPO4, r.LFe, LSi) : A::mn4(LSi, r.LN, r.LPO4, r.LFe);

    // growth rate (μ::BaseProduction)(…, L) with the SWAPPED day length — growth_rate.jl:3-47
    const double PAR = ph.blue_light_absorption * c.PAR1 + ph.green_light_absorption * c.PAR2 + ph.red_light_absorption * c.PAR3;
    const double dl = a.p.day_length_growth;
    const double dd = A::mx(0.0, c.zeu - c.zmxl);
    const double drt = A::div(dd * dd, c.kappa);
    r.mui = ph.base_growth_rate * fT;
    const double f1 = a.dv.f1_growth;
    const double f2 = 1 - A::div(drt, drt + ph.dark_tolerance);
    double alpha = ph.initial_slope_of_PI_curve;
    if (ph.low_light_adaptation != 0.0) alpha = alpha * (1 + ph.low_light_adaptation * A::ex(-PAR));
    else alpha = alpha * (1 + 0.0);
    double fl;
    if (ph.growth_rate_kind == OBM_GROWTH_NUTRIENT_LIMITED)
        fl = 1 - A::ex(A::gdiv(-alpha * r.tChl * PAR, dl * r.mui * r.L));
    else
        fl = 1 - A::ex(A::EX ? -alpha * r.tChl * PAR / (dl * (ph.basal_respiration_rate + ph.reference_growth_rate))
                           : -alpha * r.tChl * PAR * a.dv.inv_resp[cls]);
    r.mu = r.mui * f1 * f2 * fl * r.L;
    r.muI = r.mu * I;

    // mortality mixed_mondo.jl:137-167
    r.lin = A::div(ph.linear_mortality_rate * I, I + ph.mortality_half_saturation) * I;
    const double w = ph.base_quadratic_mortality
                     + A::div(ph.maximum_quadratic_mortality * 0.25 * (1 - r.L * r.L), 0.25 + r.L * r.L);
    r.quad = shear * w * (I * I);
    return r;
}

// chlorophyll synthesis: production_and_energy_assimilation_absorption_ratio (growth_rate.jl:126-156)
// + chlorophyll_growth (mixed_mondo.jl:112-124); CORRECT day-length order here
template <class A>
__device__ __forceinline__ double chlorophyll_growth(const PiscesArgs& a, const int cls, const Cell& c,
                                                     const Phyto& r, double I, double IChl) {
    const obm_pisces_phyto& ph = cls == 0 ? a.p.nano : a.p.diatoms;
    const double PAR = ph.blue_light_absorption * c.PAR1 + ph.green_light_absorption * c.PAR2 + ph.red_light_absorption * c.PAR3;
    const double dl = a.p.day_length_chlorophyll;
    const double mucheck = A::EX ? r.mu / (1.5 * dl / (dl + 0.5 * DAY)) * dl : r.mu * a.dv.dl_over_f1_chl;
    double alpha = ph.initial_slope_of_PI_curve;
    if (ph.low_light_adaptation != 0.0) alpha = alpha * (1 + ph.low_light_adaptation * A::ex(-PAR));
    else alpha = alpha * (1 + 0.0);
    const double rho = A::gdiv(12 * mucheck * I, alpha * IChl * PAR) * r.L;
    const double t0 = ph.minimum_chlorophyll_ratio, t1 = ph.maximum_chlorophyll_ratio;
    return (1 - ph.exudated_fraction) * 12 * (t0 + (t1 - t0) * rho) * r.mu * I;
}

// iron_uptake mixed_mondo.jl:177-205
template <class A>
__device__ __forceinline__ double iron_uptake(const PiscesArgs& a, const int cls, const Cell& c, const Phyto& r, double I) {
    const obm_pisces_phyto& ph = cls == 0 ? a.p.nano : a.p.diatoms;
    const double I1 = A::mn(I, ph.threshold_for_size_dependency);
    const double I2 = A::mx(0.0, I - ph.threshold_for_size_dependency);
    const double K = ph.half_saturation_for_iron_uptake * A::gdiv(I1 + ph.size_ratio * I2, I1 + I2);
    const double L1 = A::gdiv(c.Fe, c.Fe + K);
    const double L2 = 4 - A::div(4.5 * r.LFe, r.LFe + 1);
    const double q = A::EX ? r.tFe / ph.maximum_iron_ratio : r.tFe * a.dv.inv_theta_Fem[cls];
    return (1 - ph.exudated_fraction) * ph.maximum_iron_ratio * L1 * L2 * A::mx(0.0, A::div(1 - q, 1.05 - q)) * r.mui * I;
}

struct Zoo {
    double tsg, avail, ge, gI, gfI, base_ff;
    double mort, lin_mort, iron_graze, iron_ff;
};

// food_quality_dependant.jl:126-220, iron_grazing.jl:2-51 — evaluated ONCE per class
template <class A, int N>
__device__ __forceinline__ Zoo zooplankton(const obm_pisces_zoo& z, const double (&food)[4], const double (&iron)[4],
                                           double I, double fT, double dO2, double flux_C, double flux_Fe) {
    Zoo r;
    const double J = z.specific_food_threshold_concentration;
    const double base = z.maximum_grazing_rate * fT;
    double total_food = food[0] * z.food_preferences[0];
    double avail = A::mx(0.0, (food[0] - J)) * z.food_preferences[0];
    double total_iron = iron[0] * z.food_preferences[0];
    double s = A::mx(0.0, (food[0] - J)) * z.food_preferences[0] * iron[0];
#pragma unroll
    for (int n = 1; n < N; n++) {
        total_food += food[n] * z.food_preferences[n];
        const double a = A::mx(0.0, (food[n] - J)) * z.food_preferences[n];
        avail += a;
        total_iron += iron[n] * z.food_preferences[n];
        s += a * iron[n];
    }
    const double clg = A::mx(0.0, avail - A::mn(avail / 2, z.food_threshold_concentration));
    r.tsg = A::div(base * clg, z.grazing_half_saturation + total_food);
    r.avail = avail;
    const double igr = A::gdiv(total_iron, z.iron_ratio * r.tsg);
    r.ge = A::mn(1.0, igr) * A::mn(z.minimum_growth_efficiency, (1 - z.non_assimilated_fraction) * igr);
    r.gI = r.tsg * I;
    r.base_ff = z.maximum_flux_feeding_rate * fT;
    r.gfI = r.base_ff * flux_C * I;
    const double cf = A::div(I, I + z.mortality_half_saturation);
    r.mort = fT * I * (z.quadratic_mortality * I + z.linear_mortality * (cf + 3 * dO2));
    r.lin_mort = fT * z.linear_mortality * (cf + 3 * dO2) * I;
    r.iron_graze = A::gdiv(s * r.tsg, avail) * I;
    r.iron_ff = r.base_ff * flux_Fe * I;
    return r;
}
// the same with the number of prey classes as a (warp-uniform) run-time value: the rolled form of the kernel evaluates micro- and
// mesozooplankton by ONE copy of this code (operation order per class unchanged: a skipped fourth prey adds nothing)
template <class A>
__device__ __forceinline__ Zoo zooplankton_n(const obm_pisces_zoo& z, const int N, const double (&food)[4], const double (&iron)[4],
                                             double I, double fT, double dO2, double flux_C, double flux_Fe) {
    Zoo r;
    const double J = z.specific_food_threshold_concentration;
    const double base = z.maximum_grazing_rate * fT;
    double total_food = food[0] * z.food_preferences[0];
    double avail = A::mx(0.0, (food[0] - J)) * z.food_preferences[0];
    double total_iron = iron[0] * z.food_preferences[0];
    double s = A::mx(0.0, (food[0] - J)) * z.food_preferences[0] * iron[0];
#pragma unroll
    for (int n = 1; n < 4; n++) {
        if (n < N) {
            total_food += food[n] * z.food_preferences[n];
            const double a = A::mx(0.0, (food[n] - J)) * z.food_preferences[n];
            avail += a;
            total_iron += iron[n] * z.food_preferences[n];
            s += a * iron[n];
        }
    }
    const double clg = A::mx(0.0, avail - A::mn(avail / 2, z.food_threshold_concentration));
    r.tsg = A::div(base * clg, z.grazing_half_saturation + total_food);
    r.avail = avail;
    const double igr = A::gdiv(total_iron, z.iron_ratio * r.tsg);
    r.ge = A::mn(1.0, igr) * A::mn(z.minimum_growth_efficiency, (1 - z.non_assimilated_fraction) * igr);
    r.gI = r.tsg * I;
    r.base_ff = z.maximum_flux_feeding_rate * fT;
    r.gfI = r.base_ff * flux_C * I;
    const double cf = A::div(I, I + z.mortality_half_saturation);
    r.mort = fT * I * (z.quadratic_mortality * I + z.linear_mortality * (cf + 3 * dO2));
    r.lin_mort = fT * z.linear_mortality * (cf + 3 * dO2) * I;
    r.iron_graze = A::gdiv(s * r.tsg, avail) * I;
    r.iron_ff = r.base_ff * flux_Fe * I;
    return r;
}
// grazing on one prey — food_quality_dependant.jl:226-255
template <class A>
__device__ __forceinline__ double graze_on(const obm_pisces_zoo& z, const Zoo& r, double pref, double prey, double I) {
    return A::gdiv(pref * A::mx(0.0, prey - z.specific_food_threshold_concentration) * r.tsg, r.avail) * I;
}

struct Inputs {
    double P, PChl, PFe, D, DChl, DFe, DSi, Z, M, DOC, POC, GOC, SFe, BFe, PSi, CaCO3;
    Cell c;
    double PARt, Omega, wPOC, wGOC, mlPAR;
};

// All 24 tendencies of one cell → sink.
template <bool EXACT, class SINK>
__device__ __forceinline__ void cell_tendencies(const PiscesArgs& a, const Inputs& in, SINK& sink) {
    using A = Ar<EXACT>;
    const obm_pisces_params& p = a.p;
    const Cell& c = in.c;
    const double P = in.P, PChl = in.PChl, PFe = in.PFe, D = in.D, DChl = in.DChl, DFe = in.DFe, DSi = in.DSi;
    const double Z = in.Z, M = in.M, DOC = in.DOC, POC = in.POC, GOC = in.GOC, SFe = in.SFe, BFe = in.BFe;
    const double PSi = in.PSi, CaCO3 = in.CaCO3, PARt = in.PARt, Omega = in.Omega, wPOC = in.wPOC, wGOC = in.wGOC;
    auto put = [&](int n, double t) { sink.put(n, t); };

    // ---- shared scalars --------------------------------------------------------------------------------
    const double shear = c.z < c.zmxl ? p.background_shear : p.mixed_layer_shear;
    const double dO2 = A::mn(1.0, A::mx(0.0, A::div(0.4 * (p.first_anoxia_threshold - c.O2), p.second_anoxia_threshold + c.O2)));
#if !OBM_PISCES_ROLL
    // b^T once per distinct base (A::ex(T ln b); bases are parameters, ln b is host-evaluated; `same_as`
    // is uniform, so these are uniform branches / selects on scalars — no local array)
    const int s1 = a.same_as[1], s2 = a.same_as[2], s3 = a.same_as[3], s4 = a.same_as[4], s5 = a.same_as[5];
    const double fT0 = A::ex(c.T * a.ln_base[0]);
    double fT1, fT2, fT3, fT4, fT5;
    if (s1 < 0) fT1 = A::ex(c.T * a.ln_base[1]); else fT1 = fT0;
    if (s2 < 0) fT2 = A::ex(c.T * a.ln_base[2]); else fT2 = s2 == 0 ? fT0 : fT1;
    if (s3 < 0) fT3 = A::ex(c.T * a.ln_base[3]); else fT3 = s3 == 0 ? fT0 : (s3 == 1 ? fT1 : fT2);
    if (s4 < 0) fT4 = A::ex(c.T * a.ln_base[4]); else fT4 = s4 == 0 ? fT0 : (s4 == 1 ? fT1 : (s4 == 2 ? fT2 : fT3));
    if (s5 < 0) fT5 = A::ex(c.T * a.ln_base[5]); else fT5 = s5 == 0 ? fT0 : (s5 == 1 ? fT1 : (s5 == 2 ? fT2 : (s5 == 3 ? fT3 : fT4)));
    const double fT[6] = {fT0, fT1, fT2, fT3, fT4, fT5};

    // ---- phytoplankton ------------------------------------------------------------------------------------
    const Phyto n = phytoplankton<A>(a, 0, c, P, PChl, PFe, fT[0], shear);
    const Phyto d = phytoplankton<A>(a, 1, c, D, DChl, DFe, fT[1], shear);

    // ---- zooplankton ----------------------------------------------------------------------------------------
    const double fluxPOC = POC * wPOC, fluxGOC = GOC * wGOC, fluxSFe = SFe * wPOC, fluxBFe = BFe * wGOC;
    const double tSFe = A::gdiv(SFe, POC);
    const double food[4] = {P, D, POC, Z};
    const double iron[4] = {n.tFe, d.tFe, tSFe, p.micro.iron_ratio};
    const Zoo zz = zooplankton<A, 3>(p.micro, food, iron, Z, fT[2], dO2, fluxPOC + fluxGOC, fluxSFe + fluxBFe);
    const Zoo zm = zooplankton<A, 4>(p.meso, food, iron, M, fT[3], dO2, fluxPOC + fluxGOC, fluxSFe + fluxBFe);
    // grazing(zoo::MicroAndMeso, prey) = micro + meso (micro_and_meso.jl:50-52)
    const double gP_micro = graze_on<A>(p.micro, zz, p.micro.food_preferences[0], P, Z);
    const double gP_meso = graze_on<A>(p.meso, zm, p.meso.food_preferences[0], P, M);
    const double gP = gP_micro + gP_meso;
    const double gD = graze_on<A>(p.micro, zz, p.micro.food_preferences[1], D, Z) + graze_on<A>(p.meso, zm, p.meso.food_preferences[1], D, M);
    const double gPOC = graze_on<A>(p.micro, zz, p.micro.food_preferences[2], POC, Z) + graze_on<A>(p.meso, zm, p.meso.food_preferences[2], POC, M);
    const double gZ_meso = graze_on<A>(p.meso, zm, p.meso.food_preferences[3], Z, M);

    // ---- P, D, chlorophyll, iron, silicon quotas: mixed_mondo_nano_diatoms.jl:45-112 -----------------
    const double deathP = (n.lin + n.quad), deathD = (d.lin + d.quad);
    put(T_P, (1 - p.nano.exudated_fraction) * n.muI - deathP - gP);
    put(T_D, (1 - p.diatoms.exudated_fraction) * d.muI - deathD - gD);
    put(T_PChl, chlorophyll_growth<A>(a, 0, c, n, P, PChl) - (deathP + gP) * n.tChl * 12);
    put(T_DChl, chlorophyll_growth<A>(a, 1, c, d, D, DChl) - (deathD + gD) * d.tChl * 12);
    const double upFe_n = iron_uptake<A>(a, 0, c, n, P), upFe_d = iron_uptake<A>(a, 1, c, d, D);
    put(T_PFe, upFe_n - (deathP + gP) * n.tFe);
    put(T_DFe, upFe_d - (deathD + gD) * d.tFe);
#else
    // ROLLED form (-DOBM_PISCES_ROLL=1): the class-symmetric sub-models — nano / diatoms, micro / meso — and the six b^T run
    // through ONE copy of their code each, in real loops over class-indexed parameters.  Same operations in the same order
    // per class; what changes is the instruction footprint of the fast path (≈ 54 KB unrolled, against a 32 KB L1.5
    // instruction cache: ncu shows the GPC instruction cache at 91 % of its request peak, profiles/r02_full_pisces_c4.txt).
    double fTs[6];
#pragma unroll 1
    for (int u = 0; u < 6; u++) {
        const int from = a.same_as[u];
        fTs[u] = from < 0 ? A::ex(c.T * a.ln_base[u]) : fTs[from < 0 ? 0 : from];
    }
    const double fT[6] = {fTs[0], fTs[1], fTs[2], fTs[3], fTs[4], fTs[5]};

    // ---- phytoplankton ------------------------------------------------------------------------------------
    Phyto n, d;
#pragma unroll 1
    for (int cls = 0; cls < 2; cls++) {
        const Phyto r = phytoplankton<A>(a, cls, c, cls ? D : P, cls ? DChl : PChl, cls ? DFe : PFe, cls ? fT[1] : fT[0], shear);
        if (cls) d = r; else n = r;
    }

    // ---- zooplankton ----------------------------------------------------------------------------------------
    const double fluxPOC = POC * wPOC, fluxGOC = GOC * wGOC, fluxSFe = SFe * wPOC, fluxBFe = BFe * wGOC;
    const double tSFe = A::gdiv(SFe, POC);
    const double food[4] = {P, D, POC, Z};
    const double iron[4] = {n.tFe, d.tFe, tSFe, p.micro.iron_ratio};
    Zoo zz, zm;
    double gP_micro, gP_meso, gD_micro, gD_meso, gPOC_micro, gPOC_meso, gZ_meso;
#pragma unroll 1
    for (int q = 0; q < 2; q++) {
        const obm_pisces_zoo& z = q ? p.meso : p.micro;
        const double I = q ? M : Z;
        const Zoo r = zooplankton_n<A>(z, q ? 4 : 3, food, iron, I, q ? fT[3] : fT[2], dO2, fluxPOC + fluxGOC, fluxSFe + fluxBFe);
        const double g0 = graze_on<A>(z, r, z.food_preferences[0], P, I);
        const double g1 = graze_on<A>(z, r, z.food_preferences[1], D, I);
        const double g2 = graze_on<A>(z, r, z.food_preferences[2], POC, I);
        const double g3 = graze_on<A>(z, r, z.food_preferences[3], Z, I);  // used for meso only (micro_and_meso.jl:36-48)
        if (q) { zm = r; gP_meso = g0; gD_meso = g1; gPOC_meso = g2; gZ_meso = g3; }
        else { zz = r; gP_micro = g0; gD_micro = g1; gPOC_micro = g2; }
    }
    // grazing(zoo::MicroAndMeso, prey) = micro + meso (micro_and_meso.jl:50-52)
    const double gP = gP_micro + gP_meso;
    const double gD = gD_micro + gD_meso;
    const double gPOC = gPOC_micro + gPOC_meso;

    // ---- P, D, chlorophyll, iron, silicon quotas: mixed_mondo_nano_diatoms.jl:45-112 -----------------
    const double deathP = (n.lin + n.quad), deathD = (d.lin + d.quad);
    double upFe_n, upFe_d;
#pragma unroll 1
    for (int cls = 0; cls < 2; cls++) {
        const obm_pisces_phyto& ph = cls ? p.diatoms : p.nano;
        const Phyto r = cls ? d : n;
        const double I = cls ? D : P, IChl = cls ? DChl : PChl;
        const double death = cls ? deathD : deathP, g = cls ? gD : gP;
        put(cls ? T_D : T_P, (1 - ph.exudated_fraction) * r.muI - death - g);
        put(cls ? T_DChl : T_PChl, chlorophyll_growth<A>(a, cls, c, r, I, IChl) - (death + g) * r.tChl * 12);
        const double up = iron_uptake<A>(a, cls, c, r, I);
        put(cls ? T_DFe : T_PFe, up - (death + g) * r.tFe);
        if (cls) upFe_d = up; else upFe_n = up;
    }
#endif
    // silicate_uptake (diatoms) mixed_mondo.jl:217-248
    double upSi;
    {
        const obm_pisces_phyto& ph = p.diatoms;
        const double Si = c.Si, K2 = ph.enhanced_silicate_half_saturation;
        const double L1 = A::gdiv(Si, Si + ph.silicate_half_saturation);
        const double L2 = p.latitude < 0 ? A::div(Si * Si * Si, Si * Si * Si + (A::EX ? K2 * K2 * K2 : a.dv.K2_cubed)) : 0.0;
        const double F1 = A::mn4(A::gdiv(d.mu, d.mui * d.L), d.LFe, d.LPO4, d.LN);
        const double F2 = A::mn(1.0, 2.2 * A::mx(0.0, L1 - 0.5));
        const double t1 = ph.optimal_silicate_ratio * L1 * A::mn(5.4, (4.4 * A::ex(-4.23 * F1) * F2 + 1) * (1 + 2 * L2));
        upSi = (1 - ph.exudated_fraction) * t1 * d.mu * D;
    }
    const double tSi = A::gdiv(DSi, D);
    put(T_DSi, upSi - (deathD + gD) * tSi);

    // ---- Z, M: micro_and_meso.jl:36-48 -----------------------------------------------------------------------
    put(T_Z, (zz.ge * (zz.gI + zz.gfI) - zz.mort) - gZ_meso);
    put(T_M, (zm.ge * (zm.gI + zm.gfI) - zm.mort) - 0.0);

    // ---- zooplankton wastes: grazing_waste.jl, mortality_waste.jl -------------------------------------------
    const double exc_z = (1 - p.micro.non_assimilated_fraction - zz.ge) * (zz.gI + zz.gfI);
    const double exc_m = (1 - p.meso.non_assimilated_fraction - zm.ge) * (zm.gI + zm.gfI);
    const double inorg_exc = p.micro.dissolved_excretion_fraction * exc_z + p.meso.dissolved_excretion_fraction * exc_m;
    const double org_exc = (1 - p.micro.dissolved_excretion_fraction) * exc_z + (1 - p.meso.dissolved_excretion_fraction) * exc_m;
    const double ut_waste = (A::EX ? 1 / (1 - p.meso.minimum_growth_efficiency) * p.meso.quadratic_mortality : a.dv.ut_coeff) * fT[3] * (M * M);
    const double ut_R = (1 - p.meso.minimum_growth_efficiency - p.meso.non_assimilated_fraction) * ut_waste;
    const double ut_excretion = (1 - p.meso.dissolved_excretion_fraction) * ut_R;
    const double ut_respiration = p.meso.dissolved_excretion_fraction * ut_R;
    const double ut_fecal = p.meso.non_assimilated_fraction * ut_waste;

    // ---- bacteria: micro_and_meso.jl:85-132 ------------------------------------------------------------------
    const double zmin = A::mn(c.zmxl, c.zeu);
    double Bact;
    {
        const double surface = A::mn(4.0, p.microzooplankton_bacteria_concentration * Z + p.mesozooplankton_bacteria_concentration * M);
        // ifelse(z >= zₘ, 1, (zₘ / z)^a): the discarded arm has no side effect, so it is only evaluated when selected
        const double factor = c.z >= zmin ? 1.0 : pow(A::div(zmin, c.z), p.bacteria_concentration_depth_exponent);
        Bact = factor * surface;
    }
    double LBact;
    {
        const double K_NO3 = p.nitrate_half_saturation_for_bacterial_activity, K_NH4 = p.ammonia_half_saturation_for_bacterial_activity;
        const double DOC_limit = A::div(DOC, DOC + p.doc_half_saturation_for_bacterial_activity);
        const double L_N = A::div(K_NO3 * c.NH4 + K_NH4 * c.NO3, K_NO3 * K_NH4 + K_NO3 * c.NH4 + K_NH4 * c.NO3);
        const double L_PO4 = A::div(c.PO4, c.PO4 + p.phosphate_half_saturation_for_bacterial_activity);
        const double L_Fe = A::div(c.Fe, c.Fe + p.iron_half_saturation_for_bacterial_activity);
        LBact = A::mn3(L_N, L_PO4, L_Fe) * DOC_limit;
    }

    // ---- dissolved organic matter: dissolved_organic_carbon.jl:39-130 ---------------------------------------
    const double dom_deg = (A::EX ? p.dom_remineralisation_rate * fT[4] * LBact * Bact / p.dom_reference_bacteria_concentration
                                  : p.dom_remineralisation_rate * fT[4] * LBact * Bact * a.dv.inv_bact_ref) * DOC;
    const double Phi1 = shear * (p.dom_aggregation_parameters[0] * DOC + p.dom_aggregation_parameters[1] * POC) * DOC;
    const double Phi2 = shear * (p.dom_aggregation_parameters[2] * GOC) * DOC;
    const double Phi3 = (p.dom_aggregation_parameters[3] * POC + p.dom_aggregation_parameters[4] * DOC) * DOC;
    const double spec_deg = p.pom_base_breakdown_rate * fT[5] * (1 - 0.45 * dO2);  // two_size_class.jl:127-137
    put(T_DOC, ((p.nano.exudated_fraction * n.muI + p.diatoms.exudated_fraction * d.muI) + ut_excretion + org_exc + spec_deg * POC
                - dom_deg - (Phi1 + Phi2 + Phi3)));

    // ---- iron chemistry: iron/iron.jl:25-37, particulate_organic_matter/iron.jl:97-126 ------------------------
    double Fep;
    {
        const double ligands = A::mx(0.6, add_rn(mul_rn(0.09, DOC + 40), -3.0));
        const double Tk = A::mx(c.T + 273.15, 5.0);
        const double K = A::ex(add_rn(16.27, -(A::EX ? 1565.7 / Tk : div_cr(1565.7, Tk))));
        // Fe′ → Fe when ligands ≪ Fe and Fe′ → 0 when ligands ≫ Fe: either way a difference of nearly equal numbers follows
        // (here or in `colloidal` below), so the reference's rounding sequence is kept — products and sums rounded one by
        // one, a correctly rounded quotient — in the fast pass too (7 instructions more than the contracted form).
        const double Dl = add_rn(add_rn(1.0, mul_rn(K, ligands)), -mul_rn(K, c.Fe));
        const double root = sqrt(add_rn(mul_rn(Dl, Dl), mul_rn(4 * K, c.Fe)));
        Fep = A::EX ? (-Dl + root) / (2 * K) : div_cr(-Dl + root, 2 * K);
    }
    const double lFe = p.minimum_iron_scavenging_rate + p.load_specific_iron_scavenging_rate * (POC + GOC + CaCO3 + PSi);
    const double BactFe = A::div(p.maximum_bacterial_growth_rate * fT[5] * LBact * p.maximum_iron_ratio_in_bacteria * c.Fe,
                                 c.Fe + p.iron_half_saturation_for_bacteria) * Bact * p.bacterial_iron_uptake_efficiency;
    const double colloidal = 0.5 * (c.Fe - Fep);
    const double CgFe1 = A::gdiv((Phi1 + Phi3) * colloidal, DOC);
    const double CgFe2 = A::gdiv(Phi2 * colloidal, DOC);

    // ---- rain ratio & calcite: nano_diatom_coupling.jl:57-124, calcite.jl:9-19 -------------------------------
    double R;
    {
        const double L_CaCO3 = A::mn3(n.LN, A::div(c.Fe, c.Fe + 0.05), n.LPO4);
        const double pcf = A::mx(1.0, P / 2);
        const double low_light = A::div(A::mx(0.0, PARt - 1), 4 + PARt);
        const double high_light = A::div(30.0, 30 + PARt);
        const double low_T = A::mx(0.0, A::div(c.T, c.T + 0.1));
        const double high_T = 1 + A::ex(A::EX ? -((c.T - 10) * (c.T - 10)) / 25 : -((c.T - 10) * (c.T - 10)) * 0.04);
        const double depth = A::mn(1.0, A::div(-50.0, c.zmxl));
        R = (p.base_rain_ratio * L_CaCO3 * pcf * low_light * high_light * low_T * high_T * depth);
    }
    const double calcite_loss = p.micro.undissolved_calcite_fraction * gP_micro + p.meso.undissolved_calcite_fraction * gP_meso;
    const double calcite_prod = R * (calcite_loss + (n.lin + n.quad) / 2);
    double calcite_diss;
    {
        const double dCa = A::mx(0.0, 1 - Omega);
        const double e = p.calcite_dissolution_exponent;
        calcite_diss = p.base_calcite_dissolution_rate * (e == 1.0 ? dCa : pow(dCa, e)) * CaCO3;  // x^1.0 ≡ x
    }
    const double tCaCO3 = calcite_prod - calcite_diss;
    put(T_CaCO3, tCaCO3);

    // ---- POC, GOC: particulate_organic_matter/carbon.jl:3-50 --------------------------------------------------
    const double* ap = p.pom_aggregation_parameters;
    const double pom_agg = shear * (ap[0] * (POC * POC) + ap[1] * POC * GOC) + ap[2] * POC * GOC + ap[3] * (POC * POC);
    const double ff_POC = zz.base_ff * fluxPOC * Z + zm.base_ff * fluxPOC * M;  // flux_feeding(zoo, Val(:POC))
    const double ff_GOC = zz.base_ff * fluxGOC * Z + zm.base_ff * fluxGOC * M;
    const double tg_POC = gPOC + ff_POC;                                      // micro_meso_zoo_coupling.jl:27-32
    const double sm_phyto = (1 - R / 2) * (n.lin + n.quad) + d.lin / 2;         // nano_diatom_coupling.jl:1-9
    const double lm_phyto = R / 2 * (n.lin + n.quad) + d.lin / 2 + d.quad;      // :11-19
    put(T_POC, (p.micro.non_assimilated_fraction * (zz.gI + zz.gfI) + sm_phyto + zz.mort + (Phi1 + Phi3) + spec_deg * GOC
                - tg_POC - pom_agg - spec_deg * POC));
    put(T_GOC, (p.meso.non_assimilated_fraction * (zm.gI + zm.gfI) + lm_phyto + zm.lin_mort + ut_fecal + pom_agg + Phi2
                - ff_GOC - spec_deg * GOC));

    // ---- SFe, BFe: particulate_organic_matter/iron.jl:2-89 ------------------------------------------------------
    {
        const double smi = (1 - R / 2) * (n.lin + n.quad) * n.tFe + d.lin * d.tFe / 2;           // nano_diatom_coupling.jl:21-37
        const double lmi = R / 2 * (n.lin + n.quad) * n.tFe + (d.lin / 2 + d.quad) * d.tFe;       // :39-55
        const double tB = A::gdiv(BFe, GOC);
        put(T_SFe, (p.micro.non_assimilated_fraction * (zz.iron_graze + zz.iron_ff) + smi + zz.mort * p.micro.iron_ratio + spec_deg * BFe
                    + lFe * POC * Fep + p.small_fraction_of_bacterially_consumed_iron * BactFe + CgFe1
                    - tg_POC * tSFe - pom_agg * tSFe - spec_deg * SFe));
        put(T_BFe, (p.meso.non_assimilated_fraction * (zm.iron_graze + zm.iron_ff) + lmi + zm.lin_mort * p.meso.iron_ratio
                    + ut_fecal * p.meso.iron_ratio + lFe * GOC * Fep + p.large_fraction_of_bacterially_consumed_iron * BactFe + CgFe2
                    + pom_agg * tSFe - ff_GOC * tB - spec_deg * BFe));
    }

    // ---- PSi, Si: particulate_organic_matter/silicate.jl:1-48, silicate.jl:20-26 -------------------------------
    double psi_diss;
    {
        const double ll = p.fast_dissolution_rate_of_silicate, lr = p.slow_dissolution_rate_of_silicate;
        const double chi = p.base_liable_silicate_fraction * (c.z >= zmin ? 1.0 : A::ex(A::div((ll - lr) * (zmin - c.z), wGOC)));
        const double l0 = chi * ll + (1 - chi) * lr;
        const double eq = exp10(6.44 - A::div(968.0, c.T + 273.15));
        const double sat = A::div(eq - c.Si, eq);
        const double q = 1 + c.T / 400;
        const double q2 = q * q;
        const double b = (q2 * q2) * sat;  // ((1 + T/400)^4 * saturation)
        const double b2 = b * b, b4 = b2 * b2;
        const double l = l0 * (0.225 * (1 + c.T / 15) * sat + 0.775 * (b4 * b4 * b));  // (…)^9
        psi_diss = l * PSi;
    }
    put(T_PSi, (gD + d.lin + d.quad) * tSi - psi_diss);
    put(T_Si, psi_diss - upSi);

    // ---- nitrogen: nitrogen/nitrate_ammonia.jl:22-89 ------------------------------------------------------------
    const double nitrif = A::div(p.maximum_nitrification_rate * c.NH4, 1 + in.mlPAR) * (1 - dO2);
    double fixation;
    {
        const double limit = n.LN >= 0.8 ? 0.01 : 1 - n.LN;
        const double growth_requirement = A::mx(0.0, n.mui - 2.15);
        const double nutrient = A::mn(A::div(c.Fe, c.Fe + p.iron_half_saturation_for_fixation),
                                      A::div(c.PO4, c.PO4 + p.phosphate_half_saturation_for_fixation));
        const double light = 1 - A::ex(A::EX ? -PARt / p.light_saturation_for_fixation : -PARt * a.dv.inv_E);
        fixation = p.maximum_fixation_rate * growth_requirement * limit * nutrient * light;
    }
    const double upNO3 = A::gdiv(n.muI * n.LNO3, n.LN) + A::gdiv(d.muI * d.LNO3, d.LN);
    const double upNH4 = A::gdiv(n.muI * n.LNH4, n.LN) + A::gdiv(d.muI * d.LNH4, d.LN);
    const double oxic = (1 - dO2) * dom_deg, anoxic = dO2 * dom_deg;
    const double tN = p.nitrogen_redfield_ratio;
    const double tNO3 = nitrif + tN * (oxic - upNO3);
    const double tNH4 = fixation + tN * (anoxic + inorg_exc + ut_respiration - upNH4) - nitrif;
    put(T_NO3, tNO3);
    put(T_NH4, tNH4);

    // ---- PO₄, Fe, DIC, Alk, O₂ ---------------------------------------------------------------------------------------
    const double prod = n.muI + d.muI;
    put(T_PO4, p.phosphate_redfield_ratio * (inorg_exc + ut_respiration + dom_deg - prod));  // phosphate.jl:21-33
    {   // iron/simple_iron.jl:19-62
        const double Lt = p.dissolved_ligand_ratio * DOC - p.maximum_ligand_concentration;
        const double ligand_agg = p.excess_scavenging_enhancement * lFe * A::mx(0.0, c.Fe - A::mx(p.maximum_ligand_concentration, Lt)) * Fep;
        // non_assimilated_iron grazing_waste.jl:45-63, per class
        const double fz = zz.iron_graze + zz.iron_ff, fm = zm.iron_graze + zm.iron_ff;
        const double nai = (fz - p.micro.non_assimilated_fraction * fz - p.micro.iron_ratio * zz.ge * (zz.gI + zz.gfI))
                           + (fm - p.meso.non_assimilated_fraction * fm - p.meso.iron_ratio * zm.ge * (zm.gI + zm.gfI));
        put(T_Fe, (spec_deg * SFe + nai + p.meso.iron_ratio * ut_R - (upFe_n + upFe_d) - ligand_agg - (CgFe1 + CgFe2)
                   - lFe * (POC + GOC) * Fep - BactFe));
    }
    put(T_DIC, (inorg_exc + ut_respiration + dom_deg + calcite_diss - calcite_prod - prod));  // inorganic_carbon.jl:32-47
    put(T_Alk, tNH4 - tNO3 - 2 * tCaCO3);                                                   // :49-58
    {   // oxygen.jl:30-51
        const double tr = p.ratio_for_respiration, tn = p.ratio_for_nitrification;
        const double remin = ((tr + tn) * oxic + tr * anoxic);
        const double fix_c = A::EX ? tn * fixation / tN : tn * fixation * a.dv.inv_tN;
        const double nit_c = A::EX ? tn * nitrif / tN : tn * nitrif * a.dv.inv_tN;
        put(T_O2, (tr * upNH4 + (tr + tn) * upNO3 + fix_c - remin - tr * inorg_exc - tr * ut_respiration - nit_c));
    }
}


// ---- host side: kernel arguments from the C parameter block (shared by obm_pisces_tendencies and the host build) ------------
// Temperature bases are compared for equality so that b^T is evaluated once per DISTINCT base; parameter-only
// sub-expressions are evaluated once, in double precision, exactly as the device code would.
inline void pisces_prepare(PiscesArgs& A, const obm_pisces_params* p) {
    A.p = *p;
    const double bases[6] = {p->nano.temperature_sensitivity, p->diatoms.temperature_sensitivity,
                             p->micro.temperature_sensitivity, p->meso.temperature_sensitivity,
                             p->dom_temperature_sensitivity,   p->pom_temperature_sensitivity};
    for (int u = 0; u < 6; u++) {
        A.ln_base[u] = log(bases[u]);
        A.same_as[u] = -1;
        for (int q = 0; q < u; q++)
            if (bases[q] == bases[u]) { A.same_as[u] = q; break; }
    }
    PiscesArgs::Derived& dv = A.dv;
    const double dlg = p->day_length_growth, dlc = p->day_length_chlorophyll;
    dv.f1_growth = 1.5 * dlg / (dlg + 0.5 * DAY);
    dv.dl_over_f1_chl = dlc / (1.5 * dlc / (dlc + 0.5 * DAY));
    const obm_pisces_phyto* cls[2] = {&p->nano, &p->diatoms};
    const double Sip = p->silicate_climatology;
    for (int q = 0; q < 2; q++) {
        dv.inv_resp[q] = 1.0 / (dlg * (cls[q]->basal_respiration_rate + cls[q]->reference_growth_rate));
        const double pk = cls[q]->silicate_half_saturation_parameter;
        dv.KSi_add[q] = 7 * (Sip * Sip) / (pk * pk + Sip * Sip);
        dv.inv_theta_o[q] = 1.0 / cls[q]->optimal_iron_quota;
        dv.inv_theta_Fem[q] = 1.0 / cls[q]->maximum_iron_ratio;
    }
    dv.inv_bact_ref = 1.0 / p->dom_reference_bacteria_concentration;
    dv.ut_coeff = 1 / (1 - p->meso.minimum_growth_efficiency) * p->meso.quadratic_mortality;
    dv.inv_E = 1.0 / p->light_saturation_for_fixation;
    dv.inv_tN = 1.0 / p->nitrogen_redfield_ratio;
    const double K2 = p->diatoms.enhanced_silicate_half_saturation;
    dv.K2_cubed = K2 * K2 * K2;
}

}  // namespace obm
