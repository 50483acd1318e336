/*
 * oracle_carbon.c — ORACLE (test infrastructure, not product): CPU restatement of the
 * carbonate-chemistry solve of OceanBioME.jl v0.17.6, INCLUDING the reference's own damped
 * Newton–Raphson solver with its exact settings (it is the algorithm timed as the CPU baseline).
 *
 * Follows:
 *   src/Models/CarbonChemistry/carbon_chemistry.jl:111-218     (driver, outputs, solve_for_H)
 *   src/Models/CarbonChemistry/alkalinity_residual.jl:18-75    (residual + derivative)
 *   src/Models/CarbonChemistry/equilibrium_constants.jl        (K0,K1,K2,KB,KW,Is,KS,KF,KP1-3,KSi,KSP, pressure)
 *   src/Models/CarbonChemistry/virial_coefficients.jl:1-18
 *   src/Models/CarbonChemistry/calcite_concentration.jl:1-80
 *   src/Utils/solvers.jl:81-131                                (DampedNewtonRaphsonSolver, bounded_λ)
 *   src/Models/seawater_density.jl:33-39 → SeawaterPolynomials.jl 0.3 `TEOS10._ρ, τ, s, ζ`
 *     (third-party, NOT in /root/reference; Project.toml:15,28 compat "0.3"): the 55-term
 *     Roquet et al. (2015) polynomial restated from the published coefficient table.
 *
 * Pinned by the reference's own goldens (tests/test_oracle_carbon.py):
 *   docstring carbon_chemistry.jl:55-62  fCO₂ = 1308.1474527899106, pHᶠ = 7.502532746463654,
 *   fCO₂(pH=7.5) = 1315.7136384737507;  test/test_gasexchange_carbon_chem.jl:113-181.
 */
#include "oracle_common.h"

/* ---------------- TEOS-10 polynomial (SeawaterPolynomials.TEOS10) ---------------- */
static const double R000 = 8.0189615746e+02, R100 = 8.6672408165e+02, R200 = -1.7864682637e+03,
                    R300 = 2.0375295546e+03, R400 = -1.2849161071e+03, R500 = 4.3227585684e+02,
                    R600 = -6.0579916612e+01, R010 = 2.6010145068e+01, R110 = -6.5281885265e+01,
                    R210 = 8.1770425108e+01, R310 = -5.6888046321e+01, R410 = 1.7681814114e+01,
                    R510 = -1.9193502195e+00, R020 = -3.7074170417e+01, R120 = 6.1548258127e+01,
                    R220 = -6.0362551501e+01, R320 = 2.9130021253e+01, R420 = -5.4723692739e+00,
                    R030 = 2.1661789529e+01, R130 = -3.3449108469e+01, R230 = 1.9717078466e+01,
                    R330 = -3.1742946532e+00, R040 = -8.3627885467e+00, R140 = 1.1311538584e+01,
                    R240 = -5.3563304045e+00, R050 = 5.4048723791e-01, R150 = 4.8169980163e-01,
                    R060 = -1.9083568888e-01, R001 = 1.9681925209e+01, R101 = -4.2549998214e+01,
                    R201 = 5.0774768218e+01, R301 = -3.0938076334e+01, R401 = 6.6051753097e+00,
                    R011 = -1.3336301113e+01, R111 = -4.4870114575e+00, R211 = 5.0042598061e+00,
                    R311 = -6.5399043664e-01, R021 = 6.7080479603e+00, R121 = 3.5063081279e+00,
                    R221 = -1.8795372996e+00, R031 = -2.4649669534e+00, R131 = -5.5077101279e-01,
                    R041 = 5.5927935970e-01, R002 = 2.0660924175e+00, R102 = -4.9527603989e+00,
                    R202 = 2.5019633244e+00, R012 = 2.0564311499e+00, R112 = -2.1311365518e-01,
                    R022 = -1.2419983026e+00, R003 = -2.3342758797e-02, R103 = -1.8507636718e-02,
                    R013 = 3.7969820455e-01;
static const double R00 = 4.6494977072e+01, R01 = -5.2099962525e+00, R02 = 2.2601900708e-01,
                    R03 = 6.4326772569e-02, R04 = 1.5616995503e-02, R05 = -1.7243708991e-03;

static double teos10_r0(double z) { return (((((R05 * z + R04) * z + R03) * z + R02) * z + R01) * z + R00) * z; }
static double teos10_rp3(double t, double s) { return R013 * t + R103 * s + R003; }
static double teos10_rp2(double t, double s) { return (R022 * t + R112 * s + R012) * t + (R202 * s + R102) * s + R002; }
static double teos10_rp1(double t, double s) {
    return (((R041 * t + R131 * s + R031) * t + (R221 * s + R121) * s + R021) * t
            + ((R311 * s + R211) * s + R111) * s + R011) * t
           + (((R401 * s + R301) * s + R201) * s + R101) * s + R001;
}
static double teos10_rp0(double t, double s) {
    return (((((R060 * t + R150 * s + R050) * t + (R240 * s + R140) * s + R040) * t
              + ((R330 * s + R230) * s + R130) * s + R030) * t
             + (((R420 * s + R320) * s + R220) * s + R120) * s + R020) * t
            + ((((R510 * s + R410) * s + R310) * s + R210) * s + R110) * s + R010) * t
           + (((((R600 * s + R500) * s + R400) * s + R300) * s + R200) * s + R100) * s + R000;
}
static double teos10_rho(double tau, double s, double zeta) {
    double rp = ((teos10_rp3(tau, s) * zeta + teos10_rp2(tau, s)) * zeta + teos10_rp1(tau, s)) * zeta + teos10_rp0(tau, s);
    return teos10_r0(zeta) + rp;
}

/* seawater_density.jl:33-39: Z = 10·Pbar; Sa = Sp; ρ = _ρ(τ(T), s(Sa), ζ(Z)) */
double orc_teos10_polynomial_approximation(double T, double Sp, double Pbar) {
    double Z = 10 * Pbar;
    double tau = T / 40.0;
    double s = sqrt((Sp + 32.0) / (40.0 * 35.16504 / 35.0));
    double zeta = -Z / 1e4;
    return teos10_rho(tau, s, zeta);
}

/* ---------------- equilibrium constants (equilibrium_constants.jl) ---------------- */
typedef struct { double a0, a1, a2, b0, b1; } pc_t;
/* :29-40; has_P = 0 ⇔ P === nothing ⇒ 1 */
static double pressure_correction(const pc_t* pc, double Tk, int has_P, double P) {
    if (!has_P) return 1.0;
    double Tc = Tk - 273.15;
    double dV = pc->a0 + pc->a1 * Tc + pc->a2 * (Tc * Tc);
    double dk = pc->b0 + pc->b1 * Tc;
    double RT = 83.14472 * Tk;
    return exp((-dV + 0.5 * dk * P) * P / RT);
}

static const pc_t PC_K1 = {-25.50, 0.1271, 0.0, -0.00308, 0.0000877};
static const pc_t PC_K2 = {-15.82, -0.0219, 0.0, 0.00113, -0.0001475};
static const pc_t PC_KB = {-29.48, 0.1622, -0.0026080, -0.00284, 0.0};
static const pc_t PC_KW = {-20.02, 0.1119, -0.001409, -0.00513, 0.0000794};
static const pc_t PC_KS = {-18.03, 0.0466, 0.000316, -0.00453, 0.00009};
static const pc_t PC_KF = {-9.78, -0.0090, -0.000942, -0.00391, 0.000054};
static const pc_t PC_KP1 = {-14.51, 0.1211, -0.000321, -0.00267, 0.0000427};
static const pc_t PC_KP2 = {-23.12, 0.1758, -0.002647, -0.00515, 0.00009};
static const pc_t PC_KP3 = {-26.57, 0.2020, -0.0030420, -0.00408, 0.0000714};
static const pc_t PC_KSP_CALCITE = {-48.76, 0.5304, -0.0, -0.01176, 0.0003692};
static const pc_t PC_KSP_ARAGONITE = {-45.96, 0.5304, -0.0, -0.01176, 0.0003692};

/* :65-80 */
double orc_K0(double T, double S) {
    const double constant = -60.2409, inverse_T = 93.4517 * 100, log_T = 23.3585, T2 = 0.0, cS = 0.023517,
                 ST = -0.023656 / 100, ST2 = 0.0047036 / (100 * 100);
    return exp(constant + inverse_T / T + log_T * (log(T) - log(100.0)) + T2 * (T * T) + (cS + ST * T + ST2 * (T * T)) * S);
}
/* :124-126 */
double orc_K1(double T, double S, int has_P, double P) {
    return pressure_correction(&PC_K1, T, has_P, P) *
           pow(10.0, 61.2172 + -3633.86 / T + -9.67770 * log(T) + 0.011555 * S + -0.0001152 * (S * S));
}
/* :170-172 */
double orc_K2(double T, double S, int has_P, double P) {
    return pressure_correction(&PC_K2, T, has_P, P) *
           pow(10.0, -25.9290 + -471.78 / T + 0.01781 * S + -0.0001122 * (S * S) + 3.16967 * log(T));
}
/* :243-250 */
double orc_KB(double T, double S, int has_P, double P) {
    double sq = sqrt(S);
    return pressure_correction(&PC_KB, T, has_P, P) *
           exp(148.0248 + (-8966.90 + -2890.53 * sq + -77.942 * S + 1.728 * pow(S, 1.5) + -0.0996 * (S * S)) / T
               + 137.1942 * sq + 1.62142 * S + (-24.4344 + -25.085 * sq + -0.2474 * S) * log(T) + 0.053105 * sq * T);
}
/* :307-313 */
double orc_KW(double T, double S, int has_P, double P) {
    return pressure_correction(&PC_KW, T, has_P, P) *
           exp(148.9652 + -13847.26 / T + -23.6521 * log(T) + (-5.977 + 118.67 / T + 1.0495 * log(T)) * sqrt(S) + -0.01615 * S);
}
/* :341 */
double orc_ionic_strength(double S) { return 19.924 * S / (1000.0 + -1.005 * S); }
/* :410-419 */
double orc_KS(double T, double S, double Is, int has_P, double P) {
    return pressure_correction(&PC_KS, T, has_P, P) *
           exp(141.328 + -4276.1 / T + -23.093 * log(T) + (324.57 + -13856.0 / T + -47.986 * log(T)) * sqrt(Is)
               + (-771.54 + 35474.0 / T + 114.723 * log(T)) * Is + -2698.0 * pow(Is, 1.5) / T + 1776.0 * (Is * Is) / T
               + log(1 + -0.001005 * S));
}
/* :481-487 */
double orc_KF(double T, double S, double Is, double KS, int has_P, double P) {
    (void)Is;
    return pressure_correction(&PC_KF, T, has_P, P) *
           exp(-9.68 + 874.0 / T + 0.111 * sqrt(S) + log(1 + 0.0 * S) + log(1 + 0.0 * S / KS));
}
/* :523-529 with the three coefficient sets :558-651 */
static double KP(const double* c, const pc_t* pc, double T, double S, int has_P, double P) {
    return pressure_correction(pc, T, has_P, P) *
           exp(c[0] + c[1] / T + c[2] * log(T) + (c[3] + c[4] / T) * sqrt(S) + (c[5] + c[6] / T) * S);
}
static const double C_KP1[7] = {115.525, -4576.752, -18.453, 0.69171, -106.736, -0.01844, -0.65643};
static const double C_KP2[7] = {172.0883, -8814.715, -27.927, 1.3566, -160.340, -0.05778, 0.37335};
static const double C_KP3[7] = {-18.141, -3070.75, 0.0, 2.81197, 17.27039, -0.09984, -44.99486};
double orc_KP1(double T, double S, int has_P, double P) { return KP(C_KP1, &PC_KP1, T, S, has_P, P); }
double orc_KP2(double T, double S, int has_P, double P) { return KP(C_KP2, &PC_KP2, T, S, has_P, P); }
double orc_KP3(double T, double S, int has_P, double P) { return KP(C_KP3, &PC_KP3, T, S, has_P, P); }
/* :706-713 */
double orc_KSi(double T, double S, double Is) {
    return exp(117.385 + -8904.2 / T + -19.334 * log(T) + (3.5913 + -458.79 / T) * sqrt(Is) + (-1.5998 + 188.74 / T) * Is
               + (0.07871 + -12.1652 / T) * (Is * Is) + log(1 + -0.001005 * S));
}
/* :754-764 with :789-850 */
static double KSP(const double* c, const pc_t* pc, double T, double S, int has_P, double P) {
    double pcorr = pressure_correction(pc, T, has_P, P);
    double lnK_therm = c[0] + c[1] * T + c[2] / T + c[3] * log10(T);
    double lnK_sea = ((c[4] + c[5] * T + c[6] / T) * sqrt(S) + c[7] * S + c[8] * pow(S, 1.5));
    return pcorr * pow(10.0, lnK_therm + lnK_sea);
}
static const double C_CALCITE[9] = {-171.9065, -0.077993, 2839.319, 71.595, -0.77712, 0.0028426, 178.34, -0.07711, 0.0041249};
static const double C_ARAGONITE[9] = {-171.945, -0.077993, 2903.293, 71.595, -0.068393, 0.0017276, 88.135, -0.10018, 0.0059415};
double orc_KSP_calcite(double T, double S, int has_P, double P) { return KSP(C_CALCITE, &PC_KSP_CALCITE, T, S, has_P, P); }
double orc_KSP_aragonite(double T, double S, int has_P, double P) { return KSP(C_ARAGONITE, &PC_KSP_ARAGONITE, T, S, has_P, P); }

/* exposed for the pressure-correction KATs of test_gasexchange_carbon_chem.jl:134-143 */
double orc_pressure_correction(int which, double Tk, double P) {
    const pc_t* t[] = {&PC_K1, &PC_K2, &PC_KB, &PC_KW, &PC_KS, &PC_KF, &PC_KP1, &PC_KP2, &PC_KP3, &PC_KSP_CALCITE, &PC_KSP_ARAGONITE};
    return pressure_correction(t[which], Tk, 1, P);
}

/* virial_coefficients.jl:1-18.  `10^-2`, `10^-5`, `10^-6` are Float64(10)^-n. */
double orc_first_virial(double Tk) {
    double a = -1636.75, b = 12.0408, c = -3.27957 * pow(10.0, -2), d = 3.16528 * pow(10.0, -5);
    return (a + b * Tk + c * (Tk * Tk) + d * (Tk * Tk * Tk)) * pow(10.0, -6);
}
double orc_cross_virial(double Tk) { return (57.7 + -0.118 * Tk) * pow(10.0, -6); }

/* ---------------- alkalinity residual (alkalinity_residual.jl) ---------------- */
typedef struct {
    double DIC, Alk, boron, sulfate, fluoride, silicate, phosphate;
    double K1, K2, KB, KW, KS, KF, KP1, KP2, KP3, KSi;
} cc_params;

static double carbonate_denom(double H, const cc_params* p) { return H * H + p->K1 * H + p->K1 * p->K2; }
static double phosphorus_denom(double H, const cc_params* p) {
    return H * H * H + p->KP1 * (H * H) + p->KP1 * p->KP2 * H + p->KP1 * p->KP2 * p->KP3;
}
static double sulfate_denom(double H, const cc_params* p) { (void)H; return 1 + p->sulfate / p->KS; }

static double alkalinity_residual(double H, const cc_params* p) {
    double bicarbonate = p->K1 * H * p->DIC / carbonate_denom(H, p);
    double carbonate = 2 * p->DIC * p->K1 * p->K2 / carbonate_denom(H, p);
    double borate = p->boron / (1 + H / p->KB);
    double hydroxide = p->KW / H;
    double hydrogen_phosphate = p->phosphate * p->KP1 * p->KP2 * H / phosphorus_denom(H, p);
    double phosphate = 2 * p->phosphate * p->KP1 * p->KP2 * p->KP3 / phosphorus_denom(H, p);
    double silicate = p->silicate / (1 + H / p->KSi);
    double free_hydrogen = -H / sulfate_denom(H, p);
    double hydrogen_sulfate = -p->sulfate / (1 + p->KS / H * sulfate_denom(H, p));
    double hydrogen_fluoride = -p->fluoride / (1 + p->KF / H);
    double phosphoric_acid = -p->phosphate * (H * H * H) / phosphorus_denom(H, p);
    return (bicarbonate + carbonate + borate + hydroxide + hydrogen_phosphate + phosphate + silicate + free_hydrogen
            + hydrogen_sulfate + hydrogen_fluoride + phosphoric_acid - p->Alk);
}

static double d_alkalinity_residual(double H, const cc_params* p) {
    double cd = carbonate_denom(H, p), pd = phosphorus_denom(H, p), sd = sulfate_denom(H, p);
    double dcd = 2 * H + p->K1;
    double dpd = 3 * (H * H) + 2 * p->KP1 * H + p->KP1 * p->KP2;
    double dsd = 0;
    double d_bicarbonate = p->K1 * p->DIC * (p->K1 * p->K2 - H * H) / (cd * cd);
    double d_carbonate = -dcd * 2 * p->DIC * p->K1 * p->K2 / (cd * cd);
    double bden = (1 + H / p->KB);
    double d_borate = -p->boron / (bden * bden) / p->KB;
    double d_hydroxide = -p->KW / (H * H);
    double d_hydrogen_phosphate = p->phosphate * p->KP1 * p->KP2 / pd - dpd * p->phosphate * p->KP1 * p->KP2 * H / (pd * pd);
    double d_phosphate = -dpd * 2 * p->phosphate * p->KP1 * p->KP2 * p->KP3 / (pd * pd);
    double sden = (1 + H / p->KSi);
    double d_silicate = -p->silicate / (sden * sden) / p->KSi;
    double d_free_hydrogen = -1 / sd;
    double hsden = (1 + p->KS / H * sd);
    double d_hydrogen_sulfate = p->sulfate / (hsden * hsden) * (p->KS / H * dsd - p->KS / (H * H) * sd);
    double hfden = (1 + p->KF / H);
    double d_hydrogen_fluoride = -p->fluoride / (hfden * hfden) * p->KF / (H * H);
    double d_phosphoric_acid = -3 * p->phosphate * (H * H) / pd - dpd * p->phosphate * (H * H * H) / (pd * pd);
    return (d_bicarbonate + d_carbonate + d_borate + d_hydroxide + d_hydrogen_phosphate + d_phosphate + d_silicate
            + d_free_hydrogen + d_hydrogen_sulfate + d_hydrogen_fluoride + d_phosphoric_acid);
}

/* ---------------- DampedNewtonRaphsonSolver (solvers.jl:81-131) ----------------
 * settings of carbon_chemistry.jl:81: bounds = (lower = 0, upper = nothing), defaults
 * max_iters = 100, atol = 10^-20, damping = 0.5, armijo_constant = 0.5, min_damping = 0.5^10. */
static double bounded_lambda_lower(double x, double B, double delta) {
    double lb = (x - B) / delta;
    lb = lb < 0 ? 1 : lb;
    return jl_min(1, lb);
}

static double damped_newton(double x0, const cc_params* params, int* n_iters, int* n_fevals) {
    const int max_iters = 100;
    const double atol = pow(10.0, -20), damping = 0.5, c = 0.5;
    double lm = 1.0;
    for (int n = 0; n < 10; n++) lm *= damping; /* damping^10 (exact in binary) */
    double x = x0;
    int N = 0, nf = 1;
    double fx = alkalinity_residual(x, params);
    while ((fabs(fx) > atol) & (N < max_iters)) {
        double delta = fx / d_alkalinity_residual(x, params);
        double lambda = bounded_lambda_lower(x, 0.0, delta);
        double fnew = alkalinity_residual(x - lambda * delta, params);
        nf++;
        while ((fabs(fnew) >= fabs(fx) - c * lambda * fx * jl_sign(fx)) & (lambda > lm)) {
            lambda *= damping;
            fnew = alkalinity_residual(x - lambda * delta, params);
            nf++;
        }
        x -= lambda * delta;
        N += 1;
        fx = fnew;
    }
    if (n_iters) *n_iters = N;
    if (n_fevals) *n_fevals = nf;
    return x;
}

/* common front half of carbon_chemistry.jl:123-155 and calcite_concentration.jl:13-46;
 * `P_default` is the value used for density when P === nothing (1 resp. 0 — SURVEY App. A bug 4) */
static double solve_H(double DIC, double T, double S, double Alk, int has_pH, double pH, int has_P, double P,
                      double silicate, double phosphate, double initial_pH_guess, double P_default,
                      cc_params* prm, double* Tk_out, double* Is_out, int* n_iters, int* n_fevals) {
    double boron = 0.000232 / 10.811 * S / 1.80655;
    double sulfate = 0.14 / 96.06 * S / 1.80655;
    double fluoride = 0.000067 / 18.9984 * S / 1.80655;

    double rho = orc_teos10_polynomial_approximation(T, S, has_P ? P : P_default);
    T += 273.15;
    Alk *= 1e-3 / rho;
    DIC *= 1e-3 / rho;
    phosphate *= 1e-3 / rho;
    silicate *= 1e-3 / rho;
    double Is = orc_ionic_strength(S);

    prm->DIC = DIC; prm->Alk = Alk; prm->boron = boron; prm->sulfate = sulfate; prm->fluoride = fluoride;
    prm->silicate = silicate; prm->phosphate = phosphate;
    prm->K1 = orc_K1(T, S, has_P, P);
    prm->K2 = orc_K2(T, S, has_P, P);
    prm->KB = orc_KB(T, S, has_P, P);
    prm->KW = orc_KW(T, S, has_P, P);
    prm->KS = orc_KS(T, S, Is, has_P, P);
    prm->KF = orc_KF(T, S, Is, prm->KS, has_P, P);
    prm->KP1 = orc_KP1(T, S, has_P, P);
    prm->KP2 = orc_KP2(T, S, has_P, P);
    prm->KP3 = orc_KP3(T, S, has_P, P);
    prm->KSi = orc_KSi(T, S, Is);
    *Tk_out = T;
    *Is_out = Is;
    if (n_iters) *n_iters = 0;
    if (n_fevals) *n_fevals = 0;
    if (has_pH) return pow(10.0, -pH); /* carbon_chemistry.jl:213 */
    return damped_newton(pow(10.0, -initial_pH_guess), prm, n_iters, n_fevals); /* :217-218 */
}

/* (p::CarbonChemistry)(; DIC, T, S, Alk, pH, P, output, silicate, phosphate, initial_pH_guess)
 * output_kind: OBM_CC_* */
double orc_carbon_chemistry(double DIC, double T, double S, double Alk, int has_pH, double pH, int has_P, double P,
                            double silicate, double phosphate, double initial_pH_guess, int output_kind,
                            int* n_iters, int* n_fevals) {
    cc_params prm;
    double Tk, Is;
    if (output_kind == OBM_CC_CO3 || output_kind == OBM_CC_OMEGA_CALCITE) {
        /* calcite_concentration.jl:1-53 */
        double H = solve_H(DIC, T, S, Alk, has_pH, pH, has_P, P, silicate, phosphate, initial_pH_guess, 0.0, &prm, &Tk, &Is,
                           n_iters, n_fevals);
        double denom1 = (H * (H + prm.K1));
        double denom2 = (1.0 + prm.K1 * prm.K2 / denom1);
        double CO3 = prm.DIC * prm.K1 * prm.K2 / denom1 / denom2;
        if (output_kind == OBM_CC_CO3) return CO3;
        /* calcite_concentration.jl:55-80 */
        double calcium = 0.0103 * S / 35;
        double KSPc = orc_KSP_calcite(T + 273.15, S, has_P, P);
        return calcium * CO3 / KSPc;
    }
    double H = solve_H(DIC, T, S, Alk, has_pH, pH, has_P, P, silicate, phosphate, initial_pH_guess, 1.0, &prm, &Tk, &Is,
                       n_iters, n_fevals);
    double K0 = orc_K0(Tk, S);
    double CO2 = prm.DIC * (H * H) / (H * H + prm.K1 * H + prm.K1 * prm.K2);
    double fCO2 = (CO2 / K0) * 1000000.0; /* convert(FT, 10^6) */
    switch (output_kind) {
        case OBM_CC_FCO2: return fCO2;
        case OBM_CC_PH_FREE: return -log10(H);
        case OBM_CC_PCO2: { /* carbon_chemistry.jl:170-193 */
            double Pp = has_P ? P : 1.0;
            Pp *= 101325.0;
            double B = orc_first_virial(Tk), d = orc_cross_virial(Tk);
            fCO2 *= 0.09807;
            double phi = 1.0;
            double xCO2 = fCO2 / (phi * Pp);
            for (int n = 0; n < 3; n++) {
                double om = (1.0 - xCO2);
                phi = exp((B + 2.0 * (om * om) * d) * Pp / (8.31446261815324 * Tk));
                xCO2 = fCO2 / (phi * Pp);
            }
            double pCO2 = fCO2 / phi;
            pCO2 /= 0.09807;
            return pCO2;
        }
        case OBM_CC_PH_TOTAL: { /* :195-200 */
            double KS = orc_KS(Tk, S, Is, has_P, P);
            double HSO4 = prm.sulfate / (1 + KS / H);
            return -log10(H + HSO4);
        }
        case OBM_CC_PH_SEAWATER: { /* :202-210 */
            double KS = orc_KS(Tk, S, Is, has_P, P);
            double HSO4 = prm.sulfate / (1 + KS / H);
            double KF = orc_KF(Tk, S, Is, KS, has_P, P);
            double HF = prm.fluoride / (1 + KF / H);
            return -log10(H + HSO4 + HF);
        }
        default: return NAN;
    }
}

/* flat sweep (validation/carbon_chemistry.jl style): nullable optional arrays */
int orc_carbon_chemistry_sweep(int64_t n, const double* T, const double* S, const double* DIC, const double* Alk,
                               const double* P_bar, const double* silicate, const double* phosphate, const double* pH,
                               double initial_pH_guess, int output_kind, double* out, int64_t* total_fevals) {
    int64_t tot = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : tot)
    for (int64_t c = 0; c < n; c++) {
        int nf = 0;
        out[c] = orc_carbon_chemistry(DIC[c], T[c], S[c], Alk ? Alk[c] : 0.0, pH != NULL, pH ? pH[c] : 0.0, P_bar != NULL,
                                      P_bar ? P_bar[c] : 0.0, silicate ? silicate[c] : 0.0, phosphate ? phosphate[c] : 0.0,
                                      initial_pH_guess, output_kind, NULL, &nf);
        tot += nf;
    }
    if (total_fevals) *total_fevals = tot;
    return 0;
}

/* PISCES/compute_calcite_saturation.jl:21-37 — gridded Ω */
int orc_calcite_saturation(const obm_grid* g, const double* T, const double* S, const double* DIC, const double* Alk,
                           const double* Si, double* Omega) {
    int i0, i1, j0, j1;
    grid_range(g, &i0, &i1, &j0, &j1);
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int k = 0; k < g->Nz; k++)
        for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
                int64_t idx = cell_index(g, i, j, k);
                double z = g->zc[k + g->Hz];
                double P = fabs(z) * 9.80665 * 1026.0 / 100000.0;
                Omega[idx] = orc_carbon_chemistry(DIC[idx], T[idx], S[idx], Alk[idx], 0, 0.0, 1, P, Si[idx], 0.0, 8.0,
                                                  OBM_CC_OMEGA_CALCITE, NULL, NULL);
            }
    return 0;
}
