// carbon_chemistry.cuh — per-cell carbonate-system solve as a branch-free, fixed-iteration
// Newton iteration in x = ln[H⁺] with a warp-uniform early exit (device functions shared by the flat sweep kernel, the gridded
// Ω kernel and the gas-exchange kernel).
//
// Replaces src/Models/CarbonChemistry/carbon_chemistry.jl:111-210 + alkalinity_residual.jl +
// equilibrium_constants.jl + calcite_concentration.jl + Utils/solvers.jl:81-131 +
// seawater_density.jl:33-39 (SeawaterPolynomials TEOS-10).
//
// Why not the reference's solver: DampedNewtonRaphsonSolver runs with atol = 1e-20, below the
// residual's ulp, so it exits only on an exactly-zero residual — bimodal 3–10 vs 100 outer
// iterations, each failed one burning a 10-step back-tracking loop (SURVEY §8 a7, App. A bug 3).
// On SIMT hardware every warp pays the slow path.  The ROOT is what must match (|ΔpH| ≤ 1e-10);
// Newton in ln H with the step clamped to ± ln 10 is globally safe over the reference's whole
// validation box and converges to the same root to ~1e-15 in ≤ 8 (physical) / 12 (robust) steps.
//
// Work sharing: log T, √S, S^1.5, Is, √Is, log(1 − 0.001005 S) are computed once for all twelve
// constants, and each pressure correction is folded into the exponent of its constant (one
// exp per constant instead of two).
#pragma once

#include "obm_common.cuh"

namespace obm {
namespace cc {

// exp / log of the solve.  OBM_CC_EXP: 0 the library exp; 1 exp_lean of obm_common.cuh (even / odd Horner, out-of-line
// library fall-back: measured r02, +6 % on the fused scaling + Ω kernel); 2 exp_horner below — plain degree-13 Horner
// (17 FP64 instructions, one constant-bank operand each), range guard as ONE integer compare, the library exp inline in
// the cold branch (no call, so no ABI register constraints on the hot path); 3 exp_table of obm_common.cuh (64-entry table +
// degree-5 polynomial: 10 FP64 instructions per exp instead of 17; −2.5 % on the fused scaling + Ω kernel, the default).
// OBM_CC_LOG: 0 library log; 1 log_lean.
#ifndef OBM_CC_EXP
#define OBM_CC_EXP 3
#endif
#ifndef OBM_CC_LOG
#define OBM_CC_LOG 1
#endif
#ifndef OBM_CC_KSP_BATCH
#define OBM_CC_KSP_BATCH 1  // the solubility product's exp in the same branch-free block as the equilibrium constants' (prepare())
#endif
#ifndef OBM_CC_LOG1M
#define OBM_CC_LOG1M 0  // 1: ln(1 − 0.001005 S) as a 12-term series for S ≤ 50 (see constants())
#endif
#ifndef OBM_CC_BATCH
#define OBM_CC_BATCH 1  // the lean exp / log of the equilibrium constants branch-free in one basic block (see constants())
#endif
__device__ __forceinline__ double cexp(double x) {
#if OBM_CC_EXP == 0
    return exp(x);
#elif OBM_CC_EXP == 1
    return exp_lean(x);
#elif OBM_CC_EXP == 2
    return exp_horner(x);
#else
    return exp_in_range(x) ? exp_table(x) : exp(x);  // 3: the 64-entry table form of obm_common.cuh
#endif
}
__device__ __forceinline__ double cexp_unguarded(double x) {  // callers test exp_in_range themselves
#if OBM_CC_EXP == 3
    return exp_table(x);
#else
    return exp_unguarded(x);
#endif
}
__device__ __forceinline__ double clog(double x) {
#if OBM_CC_LOG == 0
    return log(x);
#else
    return log_lean(x);
#endif
}

// ---- TEOS-10 55-term polynomial (Roquet et al. 2015) as used through SeawaterPolynomials 0.3 ----
__device__ __forceinline__ double teos10_rho(double T, double Sp, double Pbar) {
    const double t = T * KD(0.025);
    const double s = sqrt((Sp + 32.0) * KD(1.0 / (40.0 * 35.16504 / 35.0)));
    const double z = -(10.0 * Pbar) * KD(1e-4);
    const double r0 = (((((KD(-1.7243708991e-03) * z + KD(1.5616995503e-02)) * z + KD(6.4326772569e-02)) * z + KD(2.2601900708e-01)) * z
                        + KD(-5.2099962525e+00)) * z + KD(4.6494977072e+01)) * z;
    const double rp3 = KD(3.7969820455e-01) * t + KD(-1.8507636718e-02) * s + KD(-2.3342758797e-02);
    const double rp2 = (KD(-1.2419983026e+00) * t + KD(-2.1311365518e-01) * s + KD(2.0564311499e+00)) * t
                       + (KD(2.5019633244e+00) * s + KD(-4.9527603989e+00)) * s + KD(2.0660924175e+00);
    const double rp1 = (((KD(5.5927935970e-01) * t + KD(-5.5077101279e-01) * s + KD(-2.4649669534e+00)) * t
                         + (KD(-1.8795372996e+00) * s + KD(3.5063081279e+00)) * s + KD(6.7080479603e+00)) * t
                        + ((KD(-6.5399043664e-01) * s + KD(5.0042598061e+00)) * s + KD(-4.4870114575e+00)) * s + KD(-1.3336301113e+01)) * t
                       + (((KD(6.6051753097e+00) * s + KD(-3.0938076334e+01)) * s + KD(5.0774768218e+01)) * s + KD(-4.2549998214e+01)) * s
                       + KD(1.9681925209e+01);
    const double rp0 = (((((KD(-1.9083568888e-01) * t + KD(4.8169980163e-01) * s + KD(5.4048723791e-01)) * t
                           + (KD(-5.3563304045e+00) * s + KD(1.1311538584e+01)) * s + KD(-8.3627885467e+00)) * t
                          + ((KD(-3.1742946532e+00) * s + KD(1.9717078466e+01)) * s + KD(-3.3449108469e+01)) * s + KD(2.1661789529e+01)) * t
                         + (((KD(-5.4723692739e+00) * s + KD(2.9130021253e+01)) * s + KD(-6.0362551501e+01)) * s + KD(6.1548258127e+01)) * s
                         + KD(-3.7074170417e+01)) * t
                        + ((((KD(-1.9193502195e+00) * s + KD(1.7681814114e+01)) * s + KD(-5.6888046321e+01)) * s + KD(8.1770425108e+01)) * s
                           + KD(-6.5281885265e+01)) * s + KD(2.6010145068e+01)) * t
                       + (((((KD(-6.0579916612e+01) * s + KD(4.3227585684e+02)) * s + KD(-1.2849161071e+03)) * s + KD(2.0375295546e+03)) * s
                           + KD(-1.7864682637e+03)) * s + KD(8.6672408165e+02)) * s + KD(8.0189615746e+02);
    return r0 + (((rp3 * z + rp2) * z + rp1) * z + rp0);
}

// ---- per-level tables: everything of the Ω solve that depends on PRESSURE only -------------------------------------------------
// In the gridded kernels a block works on one z-level (blockIdx.z = k), so the pressure is uniform over its threads.  The
// TEOS-10 polynomial collapses in ζ = −10 P / 10⁴ to a bivariate one in (s, τ): C_ij = Σ_k R_ijk ζ^k, 28 coefficients
// (+ r₀(ζ) in C_00), and every pressure correction ln Pc = (−ΔV + ½ Δκ P) P / (R T) becomes (A₀ + A₁ T_c + A₂ T_c²) / T
// with A_n(P).  The first threads of the block build the 49 numbers once in shared memory; each cell then spends 27
// FMAs on the density instead of ≈ 60 and 4 FP64 operations per correction instead of 8 (7 corrections): ≈ 60 of the
// ≈ 790 FP64 instructions per cell of the fused scaling + Ω kernel, which is bound by exactly those (ncu r3a: FP64 pipe
// 62 %, issue 66 %, 6 warps per scheduler).  Same polynomial, same coefficients, different association: ρ agrees to
// 1e-16, Ω to 1e-13 with the direct form (tests/test_fused_host.py; tests/test_gpu_pisces.py against the oracle).
#ifndef OBM_CC_LEVEL
#define OBM_CC_LEVEL 1
#endif
struct LevelTables {
    double rho[28];    // C_ij in the Horner order of teos10_rho's ζ⁰ block
    double pc[7][3];   // A₀, A₁, A₂ of K1, K2, KB, KW, KS, KF, KSP(calcite)
};
static __constant__ double TEOS_R[28][4] = {  // R_ij0 … R_ij3; (i, j) = powers of (s, τ)
    {-1.9083568888e-01, 0, 0, 0},                                                   // (0,6)
    {4.8169980163e-01, 0, 0, 0},                                                    // (1,5)
    {5.4048723791e-01, 0, 0, 0},                                                    // (0,5)
    {-5.3563304045e+00, 0, 0, 0},                                                   // (2,4)
    {1.1311538584e+01, 0, 0, 0},                                                    // (1,4)
    {-8.3627885467e+00, 5.5927935970e-01, 0, 0},                                    // (0,4)
    {-3.1742946532e+00, 0, 0, 0},                                                   // (3,3)
    {1.9717078466e+01, 0, 0, 0},                                                    // (2,3)
    {-3.3449108469e+01, -5.5077101279e-01, 0, 0},                                   // (1,3)
    {2.1661789529e+01, -2.4649669534e+00, 0, 0},                                    // (0,3)
    {-5.4723692739e+00, 0, 0, 0},                                                   // (4,2)
    {2.9130021253e+01, 0, 0, 0},                                                    // (3,2)
    {-6.0362551501e+01, -1.8795372996e+00, 0, 0},                                   // (2,2)
    {6.1548258127e+01, 3.5063081279e+00, 0, 0},                                     // (1,2)
    {-3.7074170417e+01, 6.7080479603e+00, -1.2419983026e+00, 0},                    // (0,2)
    {-1.9193502195e+00, 0, 0, 0},                                                   // (5,1)
    {1.7681814114e+01, 0, 0, 0},                                                    // (4,1)
    {-5.6888046321e+01, -6.5399043664e-01, 0, 0},                                   // (3,1)
    {8.1770425108e+01, 5.0042598061e+00, 0, 0},                                     // (2,1)
    {-6.5281885265e+01, -4.4870114575e+00, -2.1311365518e-01, 0},                   // (1,1)
    {2.6010145068e+01, -1.3336301113e+01, 2.0564311499e+00, 3.7969820455e-01},      // (0,1)
    {-6.0579916612e+01, 0, 0, 0},                                                   // (6,0)
    {4.3227585684e+02, 0, 0, 0},                                                    // (5,0)
    {-1.2849161071e+03, 6.6051753097e+00, 0, 0},                                    // (4,0)
    {2.0375295546e+03, -3.0938076334e+01, 0, 0},                                    // (3,0)
    {-1.7864682637e+03, 5.0774768218e+01, 2.5019633244e+00, 0},                     // (2,0)
    {8.6672408165e+02, -4.2549998214e+01, -4.9527603989e+00, -1.8507636718e-02},    // (1,0)
    {8.0189615746e+02, 1.9681925209e+01, 2.0660924175e+00, -2.3342758797e-02}};     // (0,0)
static __constant__ double TEOS_R0[6] = {4.6494977072e+01, -5.2099962525e+00, 2.2601900708e-01, 6.4326772569e-02,
                                         1.5616995503e-02, -1.7243708991e-03};
static __constant__ double PC_TABLE[7][5] = {  // a0, a1, a2, b0, b1 — equilibrium_constants.jl pressure corrections
    {-25.50, 0.1271, 0.0, -0.00308, 0.0000877},          // K1
    {-15.82, -0.0219, 0.0, 0.00113, -0.0001475},         // K2
    {-29.48, 0.1622, -0.0026080, -0.00284, 0.0},         // KB
    {-20.02, 0.1119, -0.001409, -0.00513, 0.0000794},    // KW
    {-18.03, 0.0466, 0.000316, -0.00453, 0.00009},       // KS
    {-9.78, -0.0090, -0.000942, -0.00391, 0.000054},     // KF
    {-48.76, 0.5304, -0.0, -0.01176, 0.0003692}};        // KSP calcite
enum { LVL_K1 = 0, LVL_K2, LVL_KB, LVL_KW, LVL_KS, LVL_KF, LVL_KSP };

// entry `n` of the tables of the level at pressure P (bar); n = 0 … 27: density, 32 … 52: corrections (one per thread)
__device__ __forceinline__ void fill_level_entry(LevelTables& t, double P, int n) {
    if (n < 28) {
        const double z = -(10.0 * P) * 1e-4;
        double c = ((TEOS_R[n][3] * z + TEOS_R[n][2]) * z + TEOS_R[n][1]) * z + TEOS_R[n][0];
        if (n == 27)
            c += (((((TEOS_R0[5] * z + TEOS_R0[4]) * z + TEOS_R0[3]) * z + TEOS_R0[2]) * z + TEOS_R0[1]) * z + TEOS_R0[0]) * z;
        t.rho[n] = c;
    } else if (n >= 32 && n < 32 + 21) {
        const int which = (n - 32) / 3, q = (n - 32) - 3 * which;
        const double* a = PC_TABLE[which];
        const double PR = P * (1.0 / 83.14472);
        t.pc[which][q] = q == 0 ? (-a[0] + 0.5 * a[3] * P) * PR : (q == 1 ? (-a[1] + 0.5 * a[4] * P) * PR : -a[2] * PR);
    }
}
__device__ __forceinline__ double teos10_rho_level(double T, double Sp, const double* C) {
    const double t = T * KD(0.025);
    const double s = sqrt((Sp + 32.0) * KD(1.0 / (40.0 * 35.16504 / 35.0)));
    return (((((C[0] * t + C[1] * s + C[2]) * t + (C[3] * s + C[4]) * s + C[5]) * t + ((C[6] * s + C[7]) * s + C[8]) * s + C[9]) * t
             + (((C[10] * s + C[11]) * s + C[12]) * s + C[13]) * s + C[14]) * t
            + ((((C[15] * s + C[16]) * s + C[17]) * s + C[18]) * s + C[19]) * s + C[20]) * t
           + (((((C[21] * s + C[22]) * s + C[23]) * s + C[24]) * s + C[25]) * s + C[26]) * s + C[27];
}
// ln of a pressure correction from the level's table
__device__ __forceinline__ double ln_pc_level(const LevelTables* l, int which, double Tc, double Tc2, double invT) {
    return (l->pc[which][0] + l->pc[which][1] * Tc + l->pc[which][2] * Tc2) * invT;
}

struct PC { double a0, a1, a2, b0, b1; };
// ln of the pressure-correction factor, equilibrium_constants.jl:29-38
__device__ __forceinline__ double ln_pc(double a0, double a1, double a2, double b0, double b1, double Tc, double P,
                                        double inv_RT) {
    const double dV = a0 + a1 * Tc + a2 * (Tc * Tc);
    const double dk = b0 + b1 * Tc;
    return (-dV + 0.5 * dk * P) * P * inv_RT;
}

// exponent of KSP calcite — equilibrium_constants.jl:754-764, :789-810 (the log10(T) in a "ln K" is the reference's, :758)
template <bool HAS_P>
__device__ __forceinline__ double KSP_calcite_exponent(double T, double S, double sqS, double logT, double P, const LevelTables* lvl) {
    constexpr double LN10 = 2.302585092994045684;
    const double iT = rcp_fast(T);
    const double therm = KD(-171.9065) + KD(-0.077993) * T + KD(2839.319) * iT + KD(71.595) * (logT * KD(1.0 / LN10));
    const double sea = ((KD(-0.77712) + KD(0.0028426) * T + KD(178.34) * iT) * sqS + KD(-0.07711) * S + KD(0.0041249) * (S * sqS));
    double e = (therm + sea) * KD(LN10);
    if (HAS_P && lvl != nullptr) {
        const double Tc = T - KD(273.15);
        e += ln_pc_level(lvl, LVL_KSP, Tc, Tc * Tc, iT);
    } else if (HAS_P)
        e += ln_pc(KD(-48.76), KD(0.5304), -0.0, KD(-0.01176), KD(0.0003692), T - KD(273.15), P, iT * KD(1.0 / 83.14472));
    return e;
}

struct Constants {
    double K1, K2, KB, KW, KS, KF, KP1, KP2, KP3, KSi;
    double Tk, Is, sqrtS, logT;
    double isd, KSsd;  // H-independent sulfate terms: 1 / (1 + ST/KS), KS (1 + ST/KS)
};

// all equilibrium constants of carbon_chemistry.jl:140-149 (defaults of :66-87)
template <bool HAS_P>
__device__ __forceinline__ void constants(double Tc_in, double S, double P, bool need_phosphate, bool need_silicate,
                                          Constants& c, const LevelTables* lvl = nullptr, bool with_KSP = false,
                                          double* KSP_out = nullptr) {
    constexpr double LN10 = 2.302585092994045684;
    const double T = Tc_in + KD(273.15);
    const double invT = rcp_fast(T);
    const double sqS = sqrt(S);
    const double S15 = S * sqS;
    const double Is = KD(19.924) * S * rcp_fast(1000.0 + KD(-1.005) * S);  // :341
    const double sqIs = sqrt(Is);
    const double Is15 = Is * sqIs;
#if OBM_CC_LOG == 1 && OBM_CC_BATCH
    // both logarithms branch-free in ONE basic block (their chains interleave); a single, rarely taken branch redoes
    // them with the library when an argument is not a positive normal number
    const double argS1 = 1 + KD(-0.001005) * S;
#if OBM_CC_LOG1M
    // ln(1 − u), u = 0.001005 S ∈ [0, 0.05] for S ≤ 50: the Mercator series −(u + u²/2 + … + u¹²/12) truncates below 4·10⁻¹⁸
    // there — 12 FMAs in place of a logarithm (exponent split, reciprocal, degree-9 polynomial in s²)
    double logT = log_unguarded(T), logS1;
    {
        const double u = KD(0.001005) * S;
        double q = KD(-1.0 / 12);
        q = fma(q, u, KD(-1.0 / 11)); q = fma(q, u, KD(-1.0 / 10)); q = fma(q, u, KD(-1.0 / 9)); q = fma(q, u, KD(-1.0 / 8));
        q = fma(q, u, KD(-1.0 / 7)); q = fma(q, u, KD(-1.0 / 6)); q = fma(q, u, KD(-1.0 / 5)); q = fma(q, u, KD(-1.0 / 4));
        q = fma(q, u, KD(-1.0 / 3)); q = fma(q, u, -0.5); q = fma(q, u, -1.0);
        logS1 = q * u;
    }
    if (!(log_in_range(T) & (S >= 0.0) & (S <= 50.0))) { logT = log(T); logS1 = log(argS1); }
#else
    double logT = log_unguarded(T), logS1 = log_unguarded(argS1);
    if (!(log_in_range(T) & log_in_range(argS1))) { logT = log(T); logS1 = log(argS1); }
#endif
#else
    const double logT = clog(T);
    const double logS1 = clog(1 + KD(-0.001005) * S);
#endif
    double Tc = 0, inv_RT = 0;
    if (HAS_P) {
        Tc = T - KD(273.15);
        inv_RT = invT * KD(1.0 / 83.14472);
    }
    // K1 :124-126, K2 :170-172 (10^x)
    double e1 = KD(61.2172) + KD(-3633.86) * invT + KD(-9.67770) * logT + KD(0.011555) * S + KD(-0.0001152) * (S * S);
    double e2 = KD(-25.9290) + KD(-471.78) * invT + KD(0.01781) * S + KD(-0.0001122) * (S * S) + KD(3.16967) * logT;
    e1 *= KD(LN10);
    e2 *= KD(LN10);
    // KB :243-250
    double eB = KD(148.0248) + (KD(-8966.90) + KD(-2890.53) * sqS + KD(-77.942) * S + KD(1.728) * S15 + KD(-0.0996) * (S * S)) * invT
                + KD(137.1942) * sqS + KD(1.62142) * S + (KD(-24.4344) + KD(-25.085) * sqS + KD(-0.2474) * S) * logT + KD(0.053105) * sqS * T;
    // KW :307-313
    double eW = KD(148.9652) + KD(-13847.26) * invT + KD(-23.6521) * logT + (KD(-5.977) + KD(118.67) * invT + KD(1.0495) * logT) * sqS + KD(-0.01615) * S;
    // KS :410-419
    double eS = KD(141.328) + KD(-4276.1) * invT + KD(-23.093) * logT + (KD(324.57) + -13856.0 * invT + KD(-47.986) * logT) * sqIs
                + (KD(-771.54) + 35474.0 * invT + KD(114.723) * logT) * Is + -2698.0 * Is15 * invT + 1776.0 * (Is * Is) * invT + logS1;
    // KF :481-487 (log(1 + 0·S) terms are exactly 0)
    double eF = KD(-9.68) + 874.0 * invT + KD(0.111) * sqS;
    if (HAS_P && lvl != nullptr) {
        const double Tc2 = Tc * Tc;
        e1 += ln_pc_level(lvl, LVL_K1, Tc, Tc2, invT);
        e2 += ln_pc_level(lvl, LVL_K2, Tc, Tc2, invT);
        eB += ln_pc_level(lvl, LVL_KB, Tc, Tc2, invT);
        eW += ln_pc_level(lvl, LVL_KW, Tc, Tc2, invT);
        eS += ln_pc_level(lvl, LVL_KS, Tc, Tc2, invT);
        eF += ln_pc_level(lvl, LVL_KF, Tc, Tc2, invT);
    } else if (HAS_P) {
        e1 += ln_pc(-25.50, KD(0.1271), 0.0, KD(-0.00308), KD(0.0000877), Tc, P, inv_RT);
        e2 += ln_pc(KD(-15.82), KD(-0.0219), 0.0, KD(0.00113), KD(-0.0001475), Tc, P, inv_RT);
        eB += ln_pc(KD(-29.48), KD(0.1622), KD(-0.0026080), KD(-0.00284), 0.0, Tc, P, inv_RT);
        eW += ln_pc(KD(-20.02), KD(0.1119), KD(-0.001409), KD(-0.00513), KD(0.0000794), Tc, P, inv_RT);
        eS += ln_pc(KD(-18.03), KD(0.0466), KD(0.000316), KD(-0.00453), KD(0.00009), Tc, P, inv_RT);
        eF += ln_pc(KD(-9.78), KD(-0.0090), KD(-0.000942), KD(-0.00391), KD(0.000054), Tc, P, inv_RT);
    }
    // KSi :706-713 (no pressure correction)
    const double eSi = KD(117.385) + KD(-8904.2) * invT + KD(-19.334) * logT + (KD(3.5913) + KD(-458.79) * invT) * sqIs
                       + (KD(-1.5998) + KD(188.74) * invT) * Is + (KD(0.07871) + KD(-12.1652) * invT) * (Is * Is) + logS1;
    c.KSi = 1.0;
    // (the calcite solubility product rides in the same batch when the caller wants it: its exponent needs T, S, √S, ln T only)
    const double eK = with_KSP ? KSP_calcite_exponent<HAS_P>(T, S, sqS, logT, P, lvl) : 0.0;
#if OBM_CC_EXP >= 2 && OBM_CC_BATCH
    // The kernel is bound by the LATENCY of dependent FP64 chains at 6 – 8 warps per scheduler (ncu, r3a: issue 66 %, FP64
    // pipe 62 %, neither saturated), and a guarded exp is its own basic block: six serial Horner chains.  Branch-free in
    // one block the six chains interleave; ONE combined range test (integer compares) sends the rare out-of-range or
    // NaN exponent through the library for all six.
    c.K1 = cexp_unguarded(e1);
    c.K2 = cexp_unguarded(e2);
    c.KB = cexp_unguarded(eB);
    c.KW = cexp_unguarded(eW);
    c.KS = cexp_unguarded(eS);
    c.KF = cexp_unguarded(eF);
    if (need_silicate) c.KSi = cexp_unguarded(eSi);
    double KSPv = 0.0;
    if (with_KSP) KSPv = cexp_unguarded(eK);
    if (!(exp_in_range(e1) & exp_in_range(e2) & exp_in_range(eB) & exp_in_range(eW) & exp_in_range(eS) & exp_in_range(eF)
          & (!need_silicate | exp_in_range(eSi)) & (!with_KSP | exp_in_range(eK)))) {
        c.K1 = exp(e1); c.K2 = exp(e2); c.KB = exp(eB); c.KW = exp(eW); c.KS = exp(eS); c.KF = exp(eF);
        if (need_silicate) c.KSi = exp(eSi);
        if (with_KSP) KSPv = exp(eK);
    }
    if (with_KSP) *KSP_out = KSPv;
#else
    if (need_silicate) c.KSi = cexp(eSi);
    c.K1 = cexp(e1);
    c.K2 = cexp(e2);
    c.KB = cexp(eB);
    c.KW = cexp(eW);
    c.KS = cexp(eS);
    c.KF = cexp(eF);
    if (with_KSP) *KSP_out = cexp(eK);
#endif
    c.KP1 = c.KP2 = c.KP3 = 1.0;
    if (need_phosphate) {  // KP1-3 :523-529, :558-651
        double p1 = KD(115.525) + KD(-4576.752) * invT + KD(-18.453) * logT + (KD(0.69171) + KD(-106.736) * invT) * sqS + (KD(-0.01844) + KD(-0.65643) * invT) * S;
        double p2 = KD(172.0883) + KD(-8814.715) * invT + KD(-27.927) * logT + (KD(1.3566) + KD(-160.340) * invT) * sqS + (KD(-0.05778) + KD(0.37335) * invT) * S;
        double p3 = KD(-18.141) + -3070.75 * invT + 0.0 * logT + (KD(2.81197) + KD(17.27039) * invT) * sqS + (KD(-0.09984) + KD(-44.99486) * invT) * S;
        if (HAS_P) {
            p1 += ln_pc(KD(-14.51), KD(0.1211), KD(-0.000321), KD(-0.00267), KD(0.0000427), Tc, P, inv_RT);
            p2 += ln_pc(KD(-23.12), KD(0.1758), KD(-0.002647), KD(-0.00515), KD(0.00009), Tc, P, inv_RT);
            p3 += ln_pc(KD(-26.57), KD(0.2020), KD(-0.0030420), KD(-0.00408), KD(0.0000714), Tc, P, inv_RT);
        }
        c.KP1 = cexp(p1);
        c.KP2 = cexp(p2);
        c.KP3 = cexp(p3);
    }
    c.Tk = T;
    c.Is = Is;
    c.sqrtS = sqS;
    c.logT = logT;
}

struct Totals {  // mol/kg (already divided by density), carbon_chemistry.jl:129-134, :116-118
    double DIC, Alk, boron, sulfate, fluoride, silicate, phosphate;
};

// alkalinity_residual(H) and H·∂ₕ residual — alkalinity_residual.jl:18-75 with shared denominators
__device__ __forceinline__ void residual(double H, const Constants& c, const Totals& t, bool need_phosphate,
                                         bool need_silicate, double& f, double& Hdf) {
    const double K1K2 = c.K1 * c.K2;
    const double cd = H * H + c.K1 * H + K1K2;
    const double icd = rcp_fast(cd);
    // bicarbonate + carbonate
    f = c.K1 * t.DIC * (H + 2 * c.K2) * icd;
    double df = c.K1 * t.DIC * ((K1K2 - H * H) - 2 * c.K2 * (2 * H + c.K1)) * (icd * icd);
    // borate
    const double ib = rcp_fast(c.KB + H);
    f += t.boron * c.KB * ib;
    df -= t.boron * c.KB * (ib * ib);
    // hydroxide − free hydrogen
    const double iH = rcp_fast(H);
    const double isd = c.isd;
    f += c.KW * iH - H * isd;
    df -= c.KW * (iH * iH) + isd;
    // hydrogen sulfate: −ST·H / (H + KS·sd)
    const double KSsd = c.KSsd;
    const double ihs = rcp_fast(H + KSsd);
    f -= t.sulfate * H * ihs;
    df -= t.sulfate * KSsd * (ihs * ihs);
    // hydrogen fluoride: −FT·H / (H + KF)
    const double ihf = rcp_fast(H + c.KF);
    f -= t.fluoride * H * ihf;
    df -= t.fluoride * c.KF * (ihf * ihf);
    if (need_silicate) {
        const double isi = rcp_fast(c.KSi + H);
        f += t.silicate * c.KSi * isi;
        df -= t.silicate * c.KSi * (isi * isi);
    }
    if (need_phosphate) {
        const double k12 = c.KP1 * c.KP2, k123 = k12 * c.KP3;
        const double H2 = H * H, H3 = H2 * H;
        const double pd = H3 + c.KP1 * H2 + k12 * H + k123;
        const double dpd = 3 * H2 + 2 * c.KP1 * H + k12;
        const double ipd = rcp_fast(pd);
        const double num = k12 * H + 2 * k123 - H3;  // [HPO₄²⁻] + 2[PO₄³⁻] − [H₃PO₄] numerator
        f += t.phosphate * num * ipd;
        df += t.phosphate * ((k12 - 3 * H2) * pd - num * dpd) * (ipd * ipd);
    }
    f -= t.Alk;
    Hdf = H * df;
}

// ---- FP32 pre-solve (OBM_CC_F32PRE) ----------------------------------------------------------------------------------
// The Newton iteration only has to END in FP64: its early steps are a search.  With no stored [H⁺] the search runs in
// FP32 — the carbonate-alkalinity quadratic with borate at the reference's initial guess, then THREE Newton steps in
// x = ln[H⁺] on the same residual, written with the O(1) speciation fractions (a₀, a₁, a₂ of the carbonate system,
// r = H / (K + H) of every one-proton pair: nothing leaves the FP32 range, no cubes of 10⁻¹⁴) and with MUFU
// reciprocals / ex2 — ≈ 55 FP32-pipe instructions per step against ≈ 125 issue slots (80 on the FP64 pipe, 7 MUFU.RCP64H
// + moves, constant loads) of an FP64 step.  FP32 rounding of the residual (≈ 6·10⁻⁸ · Alk / |∂ₓ residual|) leaves the
// iterate ≈ 3·10⁻⁶ from the root in ln H; ONE FP64 Newton step from there has the error C·Δx², C = g″ / (2 g′), and C —
// needed to three digits only — comes out of the last FP32 step for a dozen more FP32 instructions.  After the
// correction what is left is O(Δx³) ≲ 10⁻¹⁵.  The FP64 step doubles as the check: a lane whose |Δx| is not below 10⁻⁵
// (a state outside sea-water conditions, an FP32 overflow) keeps its warp in the ordinary FP64 loop, which starts from
// wherever the pre-solve ended (or from the reference's initial guess if that is not a plausible [H⁺]).
// Only the starting point differs from the reference; the root is the same (tests: |ΔpH| ≤ 1e-10 stated).
#ifndef OBM_CC_F32PRE
#define OBM_CC_F32PRE 1
#endif
#ifndef OBM_CC_F32_STEPS
#define OBM_CC_F32_STEPS 3   // FP32 Newton steps of the pre-solve
#endif
#ifndef OBM_CC_F32_REFINE
#define OBM_CC_F32_REFINE 0  // 1: refine the quadratic start once (then 2 steps do; timed in profiles/r04_kernel_variants.txt)
#endif
#ifndef OBM_CC_TOL0
#define OBM_CC_TOL0 1e-5  // exit threshold of the one FP64 step after the pre-solve
#endif
__device__ __forceinline__ float rcp_f32(float x) {
#ifdef __CUDACC__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
__device__ __forceinline__ float exp_f32(float x) {
#ifdef __CUDACC__
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
#else
    return expf(x);
#endif
}
__device__ __forceinline__ float sqrt_f32(float x) {
#ifdef __CUDACC__
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}
// → starting [H⁺] for the FP64 iteration; C2 = g″ / (2 g′) at (almost) the root, in the variable ln H
__device__ __forceinline__ double presolve_f32(const Constants& c, const Totals& t, bool need_silicate, double H_init, double& C2,
                                               bool& converged_range) {
    const float K1 = (float)c.K1, K2 = (float)c.K2, KB = (float)c.KB, KW = (float)c.KW, KSsd = (float)c.KSsd, KF = (float)c.KF;
    const float KSi = (float)c.KSi, isd = (float)c.isd;
    const float DIC = (float)t.DIC, Alk = (float)t.Alk, BT = (float)t.boron, ST = (float)t.sulfate, FT = (float)t.fluoride;
    const float SiT = (float)t.silicate;
    const float Hi = (float)H_init;
    const float K1K2 = K1 * K2;
    float H;
    {
        const float AC = Alk - BT * KB * rcp_f32(KB + Hi);
        const float b = K1 * (AC - DIC);
        const float disc = b * b - 4.0f * AC * K1K2 * (AC - 2.0f * DIC);
        H = (sqrt_f32(disc) - b) * rcp_f32(2.0f * AC);
        H = (H > 1e-12f && H < 1e-3f) ? H : Hi;  // NaN (disc < 0, A_C ≤ 0 …) fails both comparisons
#if OBM_CC_F32_REFINE
        // one refinement of the quadratic with every other species at the estimate just obtained (≈ 20 instructions: a third
        // of a Newton step) brings the start from ≈ 0.07 to ≈ 0.01 pH units, so that TWO Newton steps reach the FP32 floor
        float AC2 = Alk - BT * KB * rcp_f32(KB + H) - KW * rcp_f32(H) + H * isd;
        if (need_silicate) AC2 -= SiT * KSi * rcp_f32(KSi + H);
        const float b2 = K1 * (AC2 - DIC);
        const float disc2 = b2 * b2 - 4.0f * AC2 * K1K2 * (AC2 - 2.0f * DIC);
        const float H2 = (sqrt_f32(disc2) - b2) * rcp_f32(2.0f * AC2);
        H = (H2 > 1e-12f && H2 < 1e-3f) ? H2 : H;
#endif
    }
    float c2 = 0.0f;
    constexpr int STEPS = OBM_CC_F32_STEPS;
    // every sum below is written as a chain of fused multiply-adds (the optimiser cannot turn BT·(1 − r) into one):
    // ≈ 50 FP32-pipe instructions per step.  c0 collects the H-independent part of the residual.
    const float SiTe = need_silicate ? SiT : 0.0f;
    const float c0 = (BT + SiTe) - Alk;
#pragma unroll
    for (int n = 0; n < STEPS; n++) {
        const float icd = rcp_f32(fmaf(H + K1, H, K1K2));
        const float t = H * icd;
        const float a0 = H * t, a1 = K1 * t, a2 = K1K2 * icd;   // speciation fractions of the carbonate system
        const float m = fmaf(2.0f, a0, a1);                      // mean number of protons on them
        const float s12 = fmaf(2.0f, a2, a1);                    // their alkalinity per unit DIC
        const float rB = H * rcp_f32(KB + H), rS = H * rcp_f32(H + KSsd), rF = H * rcp_f32(H + KF);
        const float rSi = need_silicate ? H * rcp_f32(KSi + H) : 0.0f;
        const float oh = KW * rcp_f32(H);
        const float qB = fmaf(-rB, rB, rB), qS = fmaf(-rS, rS, rS), qF = fmaf(-rF, rF, rF), qSi = fmaf(-rSi, rSi, rSi);  // r (1 − r)
        float g = fmaf(DIC, s12, c0);
        g = fmaf(-BT, rB, g); g = fmaf(-SiTe, rSi, g); g = fmaf(-ST, rS, g); g = fmaf(-FT, rF, g);
        g = fmaf(-H, isd, g + oh);
        float gp = DIC * fmaf(-m, s12, a1);                      // a₁(1 − m) − 2 a₂ m
        gp = fmaf(-BT, qB, gp); gp = fmaf(-SiTe, qSi, gp); gp = fmaf(-ST, qS, gp); gp = fmaf(-FT, qF, gp);
        gp = fmaf(-H, isd, gp - oh);
        const float igp = rcp_f32(gp);
        if (n == STEPS - 1) {  // g″ at the last FP32 iterate (≲ 10⁻³ from the root: C to three digits)
            const float w = 1.0f - m;
            const float v = fmaf(2.0f * a0, 1.0f + w, a1 * w);  // ∂ₓ m = 2 a₀ (2 − m) + a₁ (1 − m)
            float gpp = DIC * fmaf(-s12, v, fmaf(a1 * w, w, 2.0f * a2 * m * m));
            gpp = fmaf(-BT * qB, fmaf(-2.0f, rB, 1.0f), gpp); gpp = fmaf(-SiTe * qSi, fmaf(-2.0f, rSi, 1.0f), gpp);
            gpp = fmaf(-ST * qS, fmaf(-2.0f, rS, 1.0f), gpp); gpp = fmaf(-FT * qF, fmaf(-2.0f, rF, 1.0f), gpp);
            gpp = fmaf(-H, isd, gpp + oh);
            c2 = 0.5f * gpp * igp;
        }
        float dx = g * igp;
        dx = dx > -2.302585f ? (dx < 2.302585f ? dx : 2.302585f) : -2.302585f;  // NaN → −ln 10: the FP64 loop sorts it out
        H *= exp_f32(-dx);
    }
    const bool ok = H > 1e-13f && H < 1e-2f && fabsf(c2) < 8.0f;
    C2 = ok ? (double)c2 : 0.0;
    converged_range = ok;
    return ok ? (double)H : H_init;
}

// solve_for_H (carbon_chemistry.jl:217-218) → [H⁺]: Newton on x = ln[H⁺] carried multiplicatively (H ← H·e^(−Δx)),
// step clamped to one pH unit, at most `iterations` steps from H0.
// OBM_CC_TOL: the warp-uniform exit threshold on |Δx|.  Newton converges quadratically here with a measured constant
// |e₊| ≈ 0.30·Δx² in ln H (sea-water states; ≤ 1 over the robust box), so leaving after a step below 10⁻⁵ puts the
// iterate within ≈ 3·10⁻¹¹ in ln H — and OBM_CC_EXTRAP (below) then removes the predictable part of that: measured
// |ΔpH| ≤ 10⁻¹³ against the reference's damped Newton with every lane leaving on its own (host build; on the GPU the
// slowest lane of a warp decides, so most lanes end far below).  Thresholds between 5·10⁻⁶ and 2·10⁻⁶ trigger one more
// step for most warps: +6 … +8 % on the fused scaling + Ω kernel (profiles/r03_kernel_variants.txt).
// OBM_CC_POLYEXP: once every lane's step is below 1/8 the factor e^(−Δx) is its degree-5 Taylor polynomial (5 FMAs
// instead of an exp); the polynomial's own error, Δx⁶/720, is part of the NEXT iterate's error like the Newton remainder
// and vanishes with it — the root is unchanged.
#ifndef OBM_CC_TOL
#define OBM_CC_TOL 1e-5
#endif
#ifndef OBM_CC_POLYEXP
#define OBM_CC_POLYEXP 1
#endif
// OBM_CC_EXTRAP: what is left after the last step is predicted and removed.  Newton's error obeys e₊ = C·e² with a
// constant that two consecutive steps reveal — Δxₙ ≈ eₙ ≈ C·Δxₙ₋₁² — so the error of the final iterate, C·Δxₙ², is
// Δxₙ³ / Δxₙ₋₁² to a relative O(Δxₙ₋₁); subtracting it costs one reciprocal and five multiplications per CELL (not per
// step) and leaves ≈ 10⁻² of the exit error: with the 10⁻⁵ threshold the root is within 10⁻¹² in ln H instead of
// 3·10⁻¹¹, at 1 % of the cost of the extra Newton step a 2·10⁻⁶ threshold would trigger (+7 %, timed).  Applied only when
// both steps are inside the asymptotic regime (|Δxₙ| < |Δxₙ₋₁| < 1/8 and the prediction below 10⁻² |Δxₙ|).
#ifndef OBM_CC_EXTRAP
#define OBM_CC_EXTRAP 1
#endif
__device__ __forceinline__ double solve_H(const Constants& c, const Totals& t, bool need_phosphate, bool need_silicate,
                                          double H0, int iterations, bool presolved = false, double C2 = 0.0) {
    constexpr double LN10 = 2.302585092994045684;
    double H = H0;
    double dx = 0.0, dx_prev = 0.0;
    int steps_done = 0;
    const unsigned mask = __activemask();
#pragma unroll 1
    for (int n = 0; n < iterations; n++) {
        steps_done = n + 1;
        double f, Hdf;
        residual(H, c, t, need_phosphate, need_silicate, f, Hdf);
        dx_prev = dx;
        dx = f * rcp_fast(Hdf);
        dx = dx < -LN10 ? -LN10 : (dx > LN10 ? LN10 : dx);  // selects, not fmin/fmax: NaN must propagate
        const double adx = fabs(dx);
#if OBM_CC_POLYEXP
        if (__all_sync(mask, adx < 0.125)) {  // warp-uniform; a NaN lane sends its warp through the exp (and stays NaN)
            double q = fma(KD(-1.0 / 120), dx, KD(1.0 / 24));
            q = fma(q, dx, KD(-1.0 / 6));
            q = fma(q, dx, 0.5);
            q = fma(q, dx, -1.0);
            H = fma(H * dx, q, H);  // H (1 − Δx + Δx²/2 − Δx³/6 + Δx⁴/24 − Δx⁵/120)
        } else
#endif
            H *= cexp(-dx);
        // Warp-uniform early exit (no divergence): once every lane's step is below the threshold the quadratic
        // convergence of Newton puts this iterate within ~1e-12 of the root; NaN lanes count as converged (they stay NaN).
        // (a FIRST step has no predecessor to extrapolate with: it ends the iteration only below 10⁻⁷ — a warm start)
        // (after the FP32 pre-solve the first step is the LAST one when it is below 10⁻⁵: its error C·Δx² is removed with
        // the C the pre-solve delivered — see presolve_f32)
        const double tol0 = presolved ? KD(OBM_CC_TOL0) : KD(1e-7);
        if (__all_sync(mask, !(adx >= (n == 0 ? tol0 : KD(OBM_CC_TOL))))) {
            if (n == 0) H = fma(-H, C2 * (dx * dx), H);  // C2 = 0 without a pre-solve
            break;
        }
    }
#if OBM_CC_EXTRAP
    if (steps_done > 1) {  // warp-uniform (the exit is a vote): a single step has no predecessor to extrapolate with
        const double ap = fabs(dx_prev), an = fabs(dx);
        const double pred = (dx * dx) * dx * rcp_fast(dx_prev * dx_prev);
        const bool use = an < ap && ap < 0.125 && fabs(pred) < KD(1e-2) * an;  // false for NaN, for a single step (Δxₙ₋₁ = 0)
        H = use ? fma(-H, pred, H) : H;  // H·e^(−pred), |pred| < 10⁻⁷
    }
#endif
    return H;
}

// Where the iteration starts when there is no stored [H⁺]: the positive root of the carbonate-alkalinity quadratic
//   A_C H² + K1 (A_C − DIC) H + K1 K2 (A_C − 2 DIC) = 0,   A_C = Alk − (every other species at the previous estimate)
// first with borate alone at the reference's initial guess H_init = 10⁻⁸ (carbon_chemistry.jl:121), then OBM_CC_INIT
// more times with borate, silicate, OH⁻ and free H⁺ evaluated at the estimate just obtained.  The plain quadratic is
// 0.07 pH units off on average for sea water (0.14 at worst: borate alkalinity moves with pH) and Newton then needs four
// steps; each refinement (≈ 55 instructions, a third of a Newton step with its derivative and exponential) halves the
// distance or better.  Timed on the B200 (profiles/r03_kernel_variants.txt): one refinement + exit at 10⁻⁵ is the
// fastest combination (two refinements + 2·10⁻⁶: +4 %).  Only the starting point differs from the reference, not the root.
#ifndef OBM_CC_INIT
#define OBM_CC_INIT 1
#endif
__device__ __forceinline__ double quadratic_H(const Constants& c, const Totals& t, double AC) {
    const double b = c.K1 * (AC - t.DIC);
    const double disc = b * b - 4.0 * AC * (c.K1 * c.K2) * (AC - 2.0 * t.DIC);
    return (sqrt(disc) - b) * rcp_fast(2.0 * AC);
}
__device__ __forceinline__ double initial_H(const Constants& c, const Totals& t, bool need_silicate, double H_init) {
    // an estimate outside pH 3 … 12 (or NaN: disc < 0, A_C ≤ 0 … fail both comparisons) is not taken
    auto plausible = [](double H) { return H > KD(1e-12) && H < KD(1e-3); };
    double H0 = quadratic_H(c, t, t.Alk - t.boron * c.KB * rcp_fast(c.KB + H_init));
    H0 = plausible(H0) ? H0 : H_init;
#pragma unroll
    for (int r = 0; r < OBM_CC_INIT; r++) {
        double AC = t.Alk - t.boron * c.KB * rcp_fast(c.KB + H0) - c.KW * rcp_fast(H0) + H0 * c.isd;
        if (need_silicate) AC -= t.silicate * c.KSi * rcp_fast(c.KSi + H0);
        const double Hn = quadratic_H(c, t, AC);
        H0 = plausible(Hn) ? Hn : H0;
    }
    return H0;
}

// K0 — equilibrium_constants.jl:65-80
__device__ __forceinline__ double K0(double T, double logT, double S) {
    return cexp(KD(-60.2409) + KD(93.4517 * 100) * rcp_fast(T) + KD(23.3585) * (logT - KD(4.605170185988092)) + 0.0 * (T * T)
               + (KD(0.023517) + KD(-0.023656 / 100) * T + KD(0.0047036 / (100.0 * 100.0)) * (T * T)) * S);
}

// KSP calcite — equilibrium_constants.jl:754-764, :789-810 (the log10(T) in a "ln K" is the reference's, :758)
template <bool HAS_P>
__device__ __forceinline__ double KSP_calcite(double T, double S, double sqS, double logT, double P, const LevelTables* lvl = nullptr) {
    return cexp(KSP_calcite_exponent<HAS_P>(T, S, sqS, logT, P, lvl));
}

// Everything of the call that depends on (T, S, P) only: density, the equilibrium constants, the S-proportional totals —
// about half of the work of an Ω solve, and none of it needs DIC, Alk or Si.  The fused scaling + Ω kernel evaluates it
// while the cell's tracers are still on their way from HBM (cp.async), then calls `finish`.
struct Prepared {
    Constants c;
    Totals t;       // boron, sulfate, fluoride filled; DIC, Alk, silicate, phosphate by `finish`
    double scale;   // 1e-3 / ρ: mmol m⁻³ → mol kg⁻¹
    double KSP;     // calcite solubility product (calcite path with `with_KSP` only)
};
template <bool HAS_P>
__device__ __forceinline__ void prepare(Prepared& q, bool calcite_path, double T, double S, double P, bool has_sil, bool has_phos,
                                        const LevelTables* lvl, bool with_KSP) {
    // density: P|1 for the main call (carbon_chemistry.jl:123), P|0 for carbonate_concentration
    // (calcite_concentration.jl:13) — reproduced as found (SURVEY App. A bug 4)
    const double rho = (HAS_P && lvl != nullptr) ? teos10_rho_level(T, S, lvl->rho)
                                                 : teos10_rho(T, S, HAS_P ? P : (calcite_path ? 0.0 : 1.0));
    q.KSP = 0.0;
#if OBM_CC_KSP_BATCH
    constants<HAS_P>(T, S, P, has_phos, has_sil, q.c, lvl, with_KSP, &q.KSP);
#else
    constants<HAS_P>(T, S, P, has_phos, has_sil, q.c, lvl);
    if (with_KSP) q.KSP = KSP_calcite<HAS_P>(q.c.Tk, S, q.c.sqrtS, q.c.logT, P, lvl);
#endif
    q.scale = KD(1e-3) * rcp_fast(rho);
    q.t.boron = KD(0.000232 / 10.811) * S * KD(1.0 / 1.80655);
    q.t.sulfate = KD(0.14 / 96.06) * S * KD(1.0 / 1.80655);
    q.t.fluoride = KD(0.000067 / 18.9984) * S * KD(1.0 / 1.80655);
    const double sd = 1 + q.t.sulfate * rcp_fast(q.c.KS);
    q.c.isd = rcp_fast(sd);
    q.c.KSsd = q.c.KS * sd;
}

// The whole `(p::CarbonChemistry)(; DIC, T, S, Alk, pH, P, output, silicate, phosphate)` call, second half.
template <bool HAS_P>
__device__ __forceinline__ double finish(Prepared& q, bool with_KSP, int output_kind, double S, double DIC, double Alk, double P,
                                         bool has_sil, double silicate, bool has_phos, double phosphate, bool has_pH,
                                         double pH, double H_init, int iterations, double* H_io, const LevelTables* lvl) {
    constexpr double LN10 = 2.302585092994045684;
    Constants& c = q.c;
    Totals& t = q.t;
    const double scale = q.scale;
    t.DIC = DIC * scale;
    t.Alk = Alk * scale;
    t.phosphate = phosphate * scale;
    t.silicate = silicate * scale;

    double H;
    if (has_pH) {
        H = cexp(-pH * KD(LN10));
    } else {
        // warm start: [H⁺] kept from the previous call on this cell, if it is a plausible value (pH 2 … 13)
        double H0 = H_io ? *H_io : 0.0;
        bool presolved = false;
        double C2 = 0.0;
        if (!(H0 > KD(1e-13) && H0 < KD(1e-2))) {
#if OBM_CC_F32PRE
            if (!has_phos) {
                H0 = presolve_f32(c, t, has_sil, H_init, C2, presolved);
            } else
#endif
                H0 = initial_H(c, t, has_sil, H_init);
        }
        H = solve_H(c, t, has_phos, has_sil, H0, iterations, presolved, C2);
        if (H_io) *H_io = H;
    }

    switch (output_kind) {
        case OBM_CC_PH_FREE: return -log10(H);
        case OBM_CC_PH_TOTAL: return -log10(H + t.sulfate * rcp_fast(1 + c.KS * rcp_fast(H)));
        case OBM_CC_PH_SEAWATER: {
            const double iH = rcp_fast(H);
            return -log10(H + t.sulfate * rcp_fast(1 + c.KS * iH) + t.fluoride * rcp_fast(1 + c.KF * iH));
        }
        case OBM_CC_CO3:
        case OBM_CC_OMEGA_CALCITE: {
            const double denom1 = (H * (H + c.K1));
            const double denom2 = (1.0 + c.K1 * c.K2 * rcp_fast(denom1));
            const double CO3 = t.DIC * c.K1 * c.K2 * rcp_fast(denom1 * denom2);
            if (output_kind == OBM_CC_CO3) return CO3;
            const double calcium = KD(0.0103) * S * KD(1.0 / 35);
            return calcium * CO3 * rcp_fast(with_KSP ? q.KSP : KSP_calcite<HAS_P>(c.Tk, S, c.sqrtS, c.logT, P, lvl));
        }
        default: break;
    }
    const double CO2 = t.DIC * (H * H) * rcp_fast(H * H + c.K1 * H + c.K1 * c.K2);
    double fCO2 = (CO2 * rcp_fast(K0(c.Tk, c.logT, S))) * 1000000.0;
    if (output_kind == OBM_CC_FCO2) return fCO2;
    // pCO₂: carbon_chemistry.jl:170-193 (3 fixed-point virial iterations)
    const double Pp = (HAS_P ? P : 1.0) * 101325.0;
    const double Tk = c.Tk;
    const double B = (-1636.75 + KD(12.0408) * Tk + KD(-3.27957e-2) * (Tk * Tk) + KD(3.16528e-5) * (Tk * Tk * Tk)) * KD(1e-6);
    const double dl = (KD(57.7) + KD(-0.118) * Tk) * KD(1e-6);
    fCO2 *= KD(0.09807);
    double phi = 1.0;
    const double iPp = rcp_fast(Pp);
    const double iRT = rcp_fast(KD(8.31446261815324) * Tk);
    double x = fCO2 * iPp;
#pragma unroll
    for (int n = 0; n < 3; n++) {
        const double om = 1.0 - x;
        phi = cexp((B + 2.0 * (om * om) * dl) * Pp * iRT);
        x = fCO2 * rcp_fast(phi) * iPp;
    }
    return fCO2 * rcp_fast(phi) * KD(1.0 / 0.09807);
}

template <bool HAS_P>
__device__ __forceinline__ double solve(int output_kind, double T, double S, double DIC, double Alk, double P,
                                        bool has_sil, double silicate, bool has_phos, double phosphate, bool has_pH,
                                        double pH, double H_init, int iterations, double* H_io = nullptr,
                                        const LevelTables* lvl = nullptr) {
    const bool calcite_path = (output_kind == OBM_CC_CO3 || output_kind == OBM_CC_OMEGA_CALCITE);
    Prepared q;
    prepare<HAS_P>(q, calcite_path, T, S, P, has_sil, has_phos, lvl, false);
    return finish<HAS_P>(q, false, output_kind, S, DIC, Alk, P, has_sil, silicate, has_phos, phosphate, has_pH, pH, H_init,
                         iterations, H_io, lvl);
}

}  // namespace cc
}  // namespace obm
