// npd_tendencies.cu — fused multi-tracer tendency kernel for the Nutrients–Plankton–Detritus
// family (NPZD, LOBSTER ± Fe ± CarbonateSystem(N) ± Oxygen, four detritus choices).
//
// Replaces the per-tracer callables of
//   src/Models/AdvectedPopulations/NutrientsPlanktonDetritus/{nutrients,plankton,detritus,
//   carbonate_system,oxygen}.jl
// which Oceananigans evaluates in one compute_Gc! launch PER TRACER: here every
// tracer of a cell is read once (coalesced along x), the shared intermediates
// (phytoplankton_growth, grazing, wastes) are evaluated once in registers, and every tendency
// is written in the same pass.  HBM-bound: 144 B/cell for LOBSTER+carbonate+O₂.
//
// The arithmetic keeps the reference's operation order inside each expression so that the only
// differences from the CPU oracle are FMA contraction and libdevice vs libm transcendentals.
#include <string.h>

#include "obm_common.cuh"

namespace obm {

constexpr int MAX_REPLICATES = 8;

struct NpdArgs {
    GridDims d;
    obm_npd_params p;
    const double *NO3, *NH4, *Fe, *N, *P, *Z, *T, *D, *sPOM, *bPOM, *DOM, *sPOC, *bPOC, *DOC, *PAR;
    double *gNO3, *gNH4, *gFe, *gN, *gP, *gZ, *gD, *gsPOM, *gbPOM, *gDOM, *gsPOC, *gbPOC, *gDOC, *gO2;
    double* gDIC[MAX_REPLICATES];
    double* gAlk[MAX_REPLICATES];
    int nrep;
    int accumulate;
    // f-2, tendencies + tracer update in one launch (obm_npd_tendencies_substep): U ← U + Δt(γG + ζG⁻), G⁻ ← G with G the
    // tendency this thread has just computed (+ the forcing found in Gⁿ when `accumulate`); u* / m* are the tracer itself
    // (written) and its G⁻, NULL for an output that is not stepped
    struct Step { int on, has_zeta, store_gn; double dt, gamma, zeta; } step;
    double *uNO3, *uNH4, *uFe, *uN, *uP, *uZ, *uD, *usPOM, *ubPOM, *uDOM, *usPOC, *ubPOC, *uDOC, *uO2;
    double *mNO3, *mNH4, *mFe, *mN, *mP, *mZ, *mD, *msPOM, *mbPOM, *mDOM, *msPOC, *mbPOC, *mDOC, *mO2;
    double *uDIC[MAX_REPLICATES], *uAlk[MAX_REPLICATES], *mDIC[MAX_REPLICATES], *mAlk[MAX_REPLICATES];
    // parameter-sweep ensembles: member = horizontal column (i + Nx·j); values[v·members + member] replaces parameter which[v]
    int nvary;
    int which[OBM_NPD_MAX_VARIED];
    const double* values;
};

// The double members of obm_npd_params in declaration order = the OBM_NPD_PARAM index space of the ensemble entry point.
#define OBM_NPD_DOUBLE_PARAMS(X)                                                                                       \
    X(nitrate_half_saturation) X(ammonia_half_saturation) X(iron_half_saturation) X(nitrate_ammonia_inhibition)        \
    X(light_half_saturation) X(phytoplankton_maximum_growth_rate) X(iron_ratio) X(phytoplankton_exudation_fraction)    \
    X(ammonia_fraction_of_exudate) X(temperature_coefficient) X(phytoplankton_mortality_rate)                          \
    X(zooplankton_mortality_rate) X(zooplankton_excretion_rate) X(phytoplankton_solid_waste_fraction)                  \
    X(excretion_inorganic_fraction) X(preference_for_phytoplankton) X(maximum_grazing_rate) X(grazing_half_saturation) \
    X(zooplankton_assimilation_fraction) X(zooplankton_calcite_dissolution) X(redfield_ratio) X(carbon_calcite_ratio)  \
    X(zooplankton_gut_calcite_dissolution) X(phytoplankton_chlorophyll_ratio) X(nitrification_rate)                    \
    X(remineralisation_inorganic_fraction) X(small_remineralisation_rate) X(large_remineralisation_rate)               \
    X(dissolved_remineralisation_rate) X(small_solid_waste_fraction) X(detritus_redfield_ratio)                        \
    X(remineralisation_rate) X(small_particle_fraction) X(respiration_oxygen_nitrogen_ratio)                           \
    X(nitrification_oxygen_nitrogen_ratio)

enum NpdParamIndex {
#define X(name) NPD_PARAM_##name,
    OBM_NPD_DOUBLE_PARAMS(X)
#undef X
    NPD_PARAM_COUNT
};
#define X(name) static_assert(offsetof(obm_npd_params, name) == offsetof(obm_npd_params, nitrate_half_saturation) + 8 * NPD_PARAM_##name, \
                              "OBM_NPD_DOUBLE_PARAMS out of step with obm_npd_params: " #name);
OBM_NPD_DOUBLE_PARAMS(X)
#undef X
static_assert(sizeof(obm_npd_params) == offsetof(obm_npd_params, nitrate_half_saturation) + 8 * NPD_PARAM_COUNT,
              "obm_npd_params has a member the ensemble index space does not name");

// `which` is uniform over the launch: a uniform jump, and the block stays in registers (no dynamically indexed struct).
__device__ __forceinline__ void set_param(obm_npd_params& p, int which, double v) {
    switch (which) {
#define X(name) case NPD_PARAM_##name: p.name = v; break;
        OBM_NPD_DOUBLE_PARAMS(X)
#undef X
        default: break;
    }
}

// `accumulate` read issued up front (with the tracer loads) so that no store waits on a late, dependent load
__device__ __forceinline__ double old_value(const double* g, long long idx, int accumulate) {
    return (accumulate && g != nullptr) ? g[idx] : 0.0;
}
// One output leaves the thread: Gⁿ ← t + old — or, in the fused tendency + substep launch, the tracer update of
// src/BoxModel/timesteppers.jl:66-93 with this value as Gⁿ (same unfused operation order as rk3_substep_kernel:
// the two paths agree bit for bit) and `cache_previous_tendencies!` (:20-28).
__device__ __forceinline__ void deliver(const NpdArgs& a, double* g, double* u, double* m, long long idx, double G) {
    if (!a.step.on) {
        if (g != nullptr) g[idx] = G;
        return;
    }
    if (g != nullptr && a.step.store_gn) g[idx] = G;
    if (u == nullptr) return;
    double rhs;
    if (a.step.has_zeta) rhs = __dmul_rn(a.step.dt, __dadd_rn(__dmul_rn(a.step.gamma, G), __dmul_rn(a.step.zeta, m[idx])));
    else rhs = __dmul_rn(__dmul_rn(a.step.dt, a.step.gamma), G);
    u[idx] = __dadd_rn(u[idx], rhs);
    if (m != nullptr) m[idx] = G;
}
#define PUT(name, value, old) deliver(a, a.g##name, a.u##name, a.m##name, idx, (value) + (old))

// plankton.jl:86-90
__device__ __forceinline__ double mortality(int form, double X, double m) { return form == OBM_LINEAR ? m * X : m * (X * X); }
__device__ __forceinline__ double concentration_limit(int form, double X, double k) {
    return form == OBM_LINEAR ? X / (X + k) : (X * X) / (X * X + k * k);
}

template <int NUT, int DET>
__device__ __forceinline__ void npd_cell(const NpdArgs& a, const obm_npd_params& p, long long idx);

// Resident blocks of 256 threads the register allocation must allow.  LOBSTER + carbonates + O₂ takes 74 registers (3 blocks) left
// alone; 4 blocks (64 registers, 16 B of stack) is faster on this HBM-bound kernel — C3: 0.678 → 0.663 ms accumulating, 0.585 → 0.537 ms
// overwriting; 5 blocks (48 registers, 104 B) 0.850 ms (visit r5i).  The parameter-sweep instantiations keep a member's 35 parameters
// in registers and are left unconstrained, like the variable-Redfield detritus (six more tracers: ≈ 90 B of stack at 64 registers).
#ifndef OBM_NPD_MIN_BLOCKS
#define OBM_NPD_MIN_BLOCKS 4
#endif
template <int NUT, int DET, bool ENSEMBLE>
__global__ void __launch_bounds__(256, (ENSEMBLE || DET == OBM_DET_VARIABLE_REDFIELD) ? 1 : OBM_NPD_MIN_BLOCKS) npd_tendency_kernel(const __grid_constant__ NpdArgs a) {
    int i, j, k;
    if (!thread_cell(a.d, i, j, k)) return;
    const long long idx = cell_index(a.d, i, j, k);
    obm_npd_params member;  // dead unless ENSEMBLE
    if constexpr (ENSEMBLE) {
        member = a.p;
        const long long members = (long long)a.d.Nx * a.d.Ny, m = i + (long long)a.d.Nx * j;
#pragma unroll 1
        for (int v = 0; v < a.nvary; v++) set_param(member, a.which[v], a.values[v * members + m]);
    }
    npd_cell<NUT, DET>(a, ENSEMBLE ? member : a.p, idx);
}

// Everything a cell does: one read of its tracers, every tendency delivered (stored, accumulated or stepped) once.
// `a` may live in the kernel's parameter space (the launches above) or in shared memory (npd_box_run_kernel below).
template <int NUT, int DET>
__device__ __forceinline__ void npd_cell(const NpdArgs& a, const obm_npd_params& p, long long idx) {
    constexpr bool HAS_NA = (NUT != OBM_NUT_NUTRIENT);
    constexpr bool TWO_SIZE = (DET == OBM_DET_TWO_PARTICLE || DET == OBM_DET_VARIABLE_REDFIELD);

    // ---- one coalesced read of everything the cell needs -----------------------------------
    const double P = a.P[idx], Z = a.Z[idx], PAR = a.PAR[idx];
    double NO3 = 0, NH4 = 0, Fe = 0, N = 0, D = 0, sPOM = 0, bPOM = 0, DOM = 0, sPOC = 0, bPOC = 0, DOC = 0;
    if constexpr (HAS_NA) { NO3 = a.NO3[idx]; NH4 = a.NH4[idx]; } else { N = a.N[idx]; }
    if constexpr (NUT == OBM_NUT_NITRATE_AMMONIA_IRON) Fe = a.Fe[idx];
    if constexpr (DET == OBM_DET_DETRITUS) D = a.D[idx];
    if constexpr (TWO_SIZE) { sPOM = a.sPOM[idx]; bPOM = a.bPOM[idx]; DOM = a.DOM[idx]; }
    if constexpr (DET == OBM_DET_VARIABLE_REDFIELD) { sPOC = a.sPOC[idx]; bPOC = a.bPOC[idx]; DOC = a.DOC[idx]; }

    const int acc = a.accumulate;
    const double oP = old_value(a.gP, idx, acc), oZ = old_value(a.gZ, idx, acc), oNO3 = old_value(a.gNO3, idx, acc),
                 oNH4 = old_value(a.gNH4, idx, acc), oFe = old_value(a.gFe, idx, acc), oN = old_value(a.gN, idx, acc),
                 oD = old_value(a.gD, idx, acc), osPOM = old_value(a.gsPOM, idx, acc), obPOM = old_value(a.gbPOM, idx, acc),
                 oDOM = old_value(a.gDOM, idx, acc), osPOC = old_value(a.gsPOC, idx, acc), obPOC = old_value(a.gbPOC, idx, acc),
                 oDOC = old_value(a.gDOC, idx, acc), oO2 = old_value(a.gO2, idx, acc),
                 oDIC = a.nrep ? old_value(a.gDIC[0], idx, acc) : 0.0, oAlk = a.nrep ? old_value(a.gAlk[0], idx, acc) : 0.0;

    // ---- growth: plankton.jl:150-220 -----------------------------------------------------------
    double nl = 0, al = 0, Ln;
    if constexpr (HAS_NA) {
        nl = NO3 * exp(-p.nitrate_ammonia_inhibition * NH4) / (NO3 + p.nitrate_half_saturation);
        al = jl_max(0.0, NH4 / (p.ammonia_half_saturation + NH4));
        Ln = (nl + al) / 2;
        if constexpr (NUT == OBM_NUT_NITRATE_AMMONIA_IRON) Ln = Ln * (Fe / (p.iron_half_saturation + Fe));
    } else {
        Ln = N / (N + p.nitrate_half_saturation);
    }
    const double kPAR = p.light_half_saturation;
    const double Ll = p.light_limitation == OBM_LIGHT_MONDO ? PAR / (kPAR + PAR) : PAR / sqrt(PAR * PAR + kPAR * kPAR);
    double Lt = 1.0;
    if (p.has_temperature_coefficient) Lt = pow(p.temperature_coefficient, a.T[idx] / 10);
    const double muP = p.phytoplankton_maximum_growth_rate * Ll * Ln * Lt * P;  // plankton.jl:205

    // ---- grazing: plankton.jl:118-135, 352-397 -----------------------------------------------
    double sPc = 0.0;  // small_particulate_concentration detritus.jl:41,103,295,308
    if constexpr (TWO_SIZE) sPc = sPOM;
    if constexpr (DET == OBM_DET_DETRITUS) sPc = D * p.small_particle_fraction;
    const double pt = p.preference_for_phytoplankton;
    const double pr = pt * P / (pt * P + (1 - pt) * sPc + eps0());
    const double food = pr * P + (1 - pr) * sPc;
    const double L = concentration_limit(p.grazing_concentration_formulation, food, p.grazing_half_saturation);
    const double g = p.maximum_grazing_rate;
    const double Gtot = g * L * Z;
    const double fden = food + jl_eps(food);
    const double Gp = g * pr * L * P / fden * Z;
    const double Gd = g * (1 - pr) * L * sPc / fden * Z;

    // ---- wastes: plankton.jl:292-346 ----------------------------------------------------------
    const double nuP = mortality(p.phytoplankton_mortality_formulation, P, p.phytoplankton_mortality_rate);
    const double exc = p.zooplankton_excretion_rate;
    const double pinw = p.excretion_inorganic_fraction * exc * Z + (1 - p.phytoplankton_solid_waste_fraction) * nuP;
    const double ponw = (1 - p.ammonia_fraction_of_exudate) * p.phytoplankton_exudation_fraction * muP
                        + (1 - p.excretion_inorganic_fraction) * exc * Z;
    const double mZZ = p.zooplankton_mortality_rate * (Z * Z);
    const double sw = (1 - p.zooplankton_assimilation_fraction) * Gtot + p.phytoplankton_solid_waste_fraction * nuP + mZZ;

    // detritus_inorganic_nitrogen_waste: detritus.jl:161-172, 298-299, 311-313
    double dinw;
    const double sm = p.small_remineralisation_rate, bm = p.large_remineralisation_rate, dm = p.dissolved_remineralisation_rate;
    const double af = p.remineralisation_inorganic_fraction;
    if constexpr (TWO_SIZE) dinw = (af * (sm * sPOM + bm * bPOM) + dm * DOM);
    else if constexpr (DET == OBM_DET_DETRITUS) dinw = D * p.remineralisation_rate;
    else dinw = ponw + sw;

    // ---- plankton: plankton.jl:92-116 -----------------------------------------------------------
    PUT(P, (1 - p.phytoplankton_exudation_fraction) * muP - Gp - nuP, oP);
    PUT(Z, p.zooplankton_assimilation_fraction * Gtot - mZZ - exc * Z, oZ);

    // ---- nutrients: nutrients.jl:22-64, uptake plankton.jl:223-278 --------------------------
    double tNO3 = 0, tNH4 = 0, tN = 0, nitrif = 0;
    const double ag = p.ammonia_fraction_of_exudate * p.phytoplankton_exudation_fraction;
    if constexpr (HAS_NA) {
        nitrif = p.nitrification_rate * NH4;
        const double lim = nl + al + eps0();
        tNO3 = nitrif - muP * nl / lim;
        tNH4 = pinw + dinw - nitrif - (muP * al / lim - ag * muP);
        PUT(NO3, tNO3, oNO3);
        PUT(NH4, tNH4, oNH4);
        if constexpr (NUT == OBM_NUT_NITRATE_AMMONIA_IRON) PUT(Fe, -(p.iron_ratio * muP), oFe);
    } else {
        tN = pinw + dinw - muP * (1 - ag);
        PUT(N, tN, oN);
    }

    // ---- detritus: detritus.jl:85-101, 143-158, 282-288 --------------------------------------
    const double R = p.redfield_ratio;
    if constexpr (DET == OBM_DET_DETRITUS) {
        PUT(D, ponw + sw - Gd - p.remineralisation_rate * D, oD);
    }
    if constexpr (TWO_SIZE) {
        const double ssf = p.small_solid_waste_fraction;
        PUT(sPOM, ssf * sw - Gd - sm * sPOM, osPOM);
        PUT(bPOM, (1 - ssf) * sw - bm * bPOM, obPOM);
        PUT(DOM, ponw + (1 - af) * (sm * sPOM + bm * bPOM) - dm * DOM, oDOM);
        if constexpr (DET == OBM_DET_VARIABLE_REDFIELD) {
            const double scw = sw * R;  // solid_carbon_waste plankton.jl:348
            // calcite_production plankton.jl:399-412
            const double cprod = (Gp * (1 - p.zooplankton_gut_calcite_dissolution) + nuP) * p.carbon_calcite_ratio * R;
            PUT(sPOC, ssf * scw - Gd * R - sm * sPOC, osPOC);
            PUT(bPOC, (1 - ssf) * scw + cprod - bm * bPOC, obPOC);
            PUT(DOC, R * ponw + (1 - af) * (sm * sPOC + bm * bPOC) - dm * DOC, oDOC);
        }
    }

    // ---- carbonate system: carbonate_system.jl:50-68 ------------------------------------------
    double cdis = 0, cupt = 0;
    if (a.nrep > 0) {
        const double rho = p.carbon_calcite_ratio, gam = p.phytoplankton_exudation_fraction;
        // phytoplankton_primary_production plankton.jl:280-289
        const double ppp = (1 + rho * (1 - gam) - ag) * muP * R;
        // detritus_inorganic_carbon_waste detritus.jl:185-196, 301-302, 315-317
        double dicw;
        if constexpr (DET == OBM_DET_TWO_PARTICLE) {
            const double Rd = p.detritus_redfield_ratio;
            dicw = af * (sm * (sPOM * Rd) + bm * (bPOM * Rd)) + dm * (DOM * Rd);
        } else if constexpr (DET == OBM_DET_VARIABLE_REDFIELD) {
            dicw = af * (sm * sPOC + bm * bPOC) + dm * DOC;
        } else if constexpr (DET == OBM_DET_DETRITUS) {
            dicw = D * p.remineralisation_rate * p.detritus_redfield_ratio;
        } else {
            dicw = (ponw + sw) * R;
        }
        // calcite_dissolution plankton.jl:414-441 (fixed-Redfield override unless VariableRedfield)
        if constexpr (DET == OBM_DET_VARIABLE_REDFIELD) cdis = Gp * p.zooplankton_gut_calcite_dissolution * rho * R;
        else cdis = (Gp + nuP) * rho * R;
        cupt = 2 * rho * muP * R;  // calcite_uptake plankton.jl:443-450
        const double tDIC = -ppp + R * pinw + dicw + cdis;
        double tAlk;
        if constexpr (HAS_NA) tAlk = tNH4 * (1 - 1.0 / 16) - tNO3 * (1 + 1.0 / 16) - 2.0 * cupt + 2.0 * cdis;
        else tAlk = tN - 2.0 * cupt + 2.0 * cdis;
        deliver(a, a.gDIC[0], a.uDIC[0], a.mDIC[0], idx, tDIC + oDIC);
        deliver(a, a.gAlk[0], a.uAlk[0], a.mAlk[0], idx, tAlk + oAlk);
#pragma unroll 1
        for (int r = 1; r < a.nrep; r++) {  // CarbonateSystem(N): every replicate gets the same tendency (:70-83)
            deliver(a, a.gDIC[r], a.uDIC[r], a.mDIC[r], idx, tDIC + old_value(a.gDIC[r], idx, a.accumulate));
            deliver(a, a.gAlk[r], a.uAlk[r], a.mAlk[r], idx, tAlk + old_value(a.gAlk[r], idx, a.accumulate));
        }
    }

    // ---- oxygen: oxygen.jl:21-31 (Nutrient models: bgc(Val(:NH₄)) = 0, nitrification = 0) -------
    if (a.gO2 != nullptr || a.uO2 != nullptr) {
        const double Rp = p.respiration_oxygen_nitrogen_ratio, Rn = p.nitrification_oxygen_nitrogen_ratio;
        PUT(O2, Rp * muP - (Rp - Rn) * tNH4 - Rp * nitrif, oO2);
    }
}

// ---- the whole RUN of a box-model ensemble in one launch (obm_npd_box_run) ----------------------------------------------------
// Boxes do not talk to each other, so nothing forces a launch per stage: a thread owns a box and integrates it through every
// stage of every time step.  The stage itself is npd_cell — the code of the fused tendency + substep launch, unchanged —
// run on a copy of the kernel arguments in shared memory whose field pointers are redirected to the block's staging
// arrays (one slot per thread and field: tracers, G⁻, PAR, T), so that a stage costs shared-memory traffic and arithmetic
// only; thread 0 rewrites the stage's (γ, ζ) between two block barriers.  The prescribed series (PAR, T: one value per
// stage, shared or per box) come from tables the host tabulated, exactly as the CUDA-graph path reads them; snapshots go
// straight to their device arrays.  Same operations in the same order as a replayed graph of obm_npd_tendencies_substep
// launches ⇒ the same bits (tests/test_gpu_box_model.py); 1000 RK3 steps of the reference's NPZD box benchmark
// (benchmark/box_model.jl: 23.5 ms on its CPU) take one launch instead of ≈ 12 000.
constexpr int RUN_BLOCK = 64;
constexpr int RUN_MAX_SLOTS = OBM_NPD_MAX_TRACERS + 2 * MAX_REPLICATES;
struct RunArgs {
    int nsteps, nstages, ncells, output_every, nslots;
    double gamma[3], zeta[3];
    int has_zeta[3];
    double* PAR_field;              // written back at the end: the state update's last prescribed value
    const double* PAR_table;        // [nsteps·nstages][PAR_per_box ? ncells : 1]; row r = what the stage AFTER global stage r sees
    int PAR_per_box;
    double* T_field;                // nullable (no temperature dependence)
    const double* T_table;          // nullable: T is then constant over the run
    int T_per_box;
    double* snap[RUN_MAX_SLOTS];    // per staged tracer (slot order): [nsteps / output_every][ncells], or nullptr
};

template <int NUT, int DET, bool ENSEMBLE>
__global__ void __launch_bounds__(RUN_BLOCK) npd_box_run_kernel(const __grid_constant__ NpdArgs a, const __grid_constant__ RunArgs r) {
    __shared__ NpdArgs sa;
    __shared__ double* gU[RUN_MAX_SLOTS];
    __shared__ double* gM[RUN_MAX_SLOTS];
    extern __shared__ double stage_mem[];  // U[nslots][RUN_BLOCK], M[nslots][RUN_BLOCK], PAR[RUN_BLOCK], T[RUN_BLOCK]
    double* U = stage_mem;
    double* M = U + (size_t)r.nslots * RUN_BLOCK;
    double* PARs = M + (size_t)r.nslots * RUN_BLOCK;
    double* Ts = PARs + RUN_BLOCK;
    const int tid = threadIdx.x;
    static_assert(sizeof(NpdArgs) % sizeof(double) == 0, "NpdArgs is copied in 8-byte words");
    for (int w = tid; w < (int)(sizeof(NpdArgs) / sizeof(double)); w += RUN_BLOCK)
        reinterpret_cast<double*>(&sa)[w] = reinterpret_cast<const double*>(&a)[w];
    __syncthreads();
    if (tid == 0) {
        int s = 0;
#define OBM_STAGE(name)                                                   \
    if (a.u##name != nullptr) {                                           \
        gU[s] = a.u##name; gM[s] = a.m##name;                             \
        sa.name = sa.u##name = U + s * RUN_BLOCK; sa.m##name = M + s * RUN_BLOCK; s++; \
    }
        OBM_STAGE(NO3) OBM_STAGE(NH4) OBM_STAGE(Fe) OBM_STAGE(N) OBM_STAGE(P) OBM_STAGE(Z) OBM_STAGE(D) OBM_STAGE(sPOM)
        OBM_STAGE(bPOM) OBM_STAGE(DOM) OBM_STAGE(sPOC) OBM_STAGE(bPOC) OBM_STAGE(DOC)
#undef OBM_STAGE
        for (int q = 0; q < a.nrep; q++) {  // one-way coupled: stepped, never read
            gU[s] = a.uDIC[q]; gM[s] = a.mDIC[q]; sa.uDIC[q] = U + s * RUN_BLOCK; sa.mDIC[q] = M + s * RUN_BLOCK; s++;
            gU[s] = a.uAlk[q]; gM[s] = a.mAlk[q]; sa.uAlk[q] = U + s * RUN_BLOCK; sa.mAlk[q] = M + s * RUN_BLOCK; s++;
        }
        if (a.uO2 != nullptr) { gU[s] = a.uO2; gM[s] = a.mO2; sa.uO2 = U + s * RUN_BLOCK; sa.mO2 = M + s * RUN_BLOCK; s++; }
        sa.PAR = PARs;
        if (a.T != nullptr) sa.T = Ts;
    }
    __syncthreads();
    const int cell = blockIdx.x * RUN_BLOCK + tid;
    const bool inside = cell < r.ncells;
    const int nx = a.d.Nx, ibox = a.d.i0 + (inside ? cell : 0);  // tables, parameter values and snapshots are indexed by the box
    const long long gidx = cell_index(a.d, ibox, a.d.j0, 0);
    obm_npd_params member;  // dead unless ENSEMBLE
    if constexpr (ENSEMBLE) {
        member = a.p;
        if (inside) {
#pragma unroll 1
            for (int v = 0; v < a.nvary; v++) set_param(member, a.which[v], a.values[(long long)v * nx + ibox]);
        }
    }
    for (int s = 0; s < r.nslots; s++) {
        U[s * RUN_BLOCK + tid] = inside ? gU[s][gidx] : 0.0;
        M[s * RUN_BLOCK + tid] = inside ? gM[s][gidx] : 0.0;
    }
    double PARv = inside ? r.PAR_field[gidx] : 0.0;
    double Tv = (inside && r.T_field != nullptr) ? r.T_field[gidx] : 0.0;
    long long row = 0;
    for (int it = 0; it < r.nsteps; it++) {
#pragma unroll 1
        for (int st = 0; st < r.nstages; st++) {
            if (tid == 0) {
                sa.step.gamma = r.gamma[st];
                sa.step.zeta = r.zeta[st];
                sa.step.has_zeta = r.has_zeta[st];
            }
            PARs[tid] = PARv;
            Ts[tid] = Tv;
            __syncthreads();
            if constexpr (ENSEMBLE) {
                // the member's parameters are opaque to the optimiser at every stage, as they are to a per-stage launch:
                // a parameter-only product hoisted out of this loop would be rounded on its own where the per-stage
                // kernel contracts it into an FMA — the two paths would drift apart by an ulp
#define X(name) asm volatile("" : "+d"(member.name));
                OBM_NPD_DOUBLE_PARAMS(X)
#undef X
            }
            if (inside) npd_cell<NUT, DET>(sa, ENSEMBLE ? member : sa.p, tid);
            __syncthreads();  // every thread has read this stage's coefficients
            // what update_state! leaves for the next stage: the prescribed series at the time after this one
            if (inside) {
                PARv = r.PAR_table[r.PAR_per_box ? row * nx + ibox : row];
                if (r.T_table != nullptr) Tv = r.T_table[r.T_per_box ? row * nx + ibox : row];
            }
            row++;
        }
        if (r.output_every > 0 && (it + 1) % r.output_every == 0 && inside) {
            const long long o = (long long)((it + 1) / r.output_every - 1) * nx + ibox;
            for (int s = 0; s < r.nslots; s++)
                if (r.snap[s] != nullptr) r.snap[s][o] = U[s * RUN_BLOCK + tid];
        }
    }
    if (!inside) return;
    for (int s = 0; s < r.nslots; s++) {
        gU[s][gidx] = U[s * RUN_BLOCK + tid];
        gM[s][gidx] = M[s * RUN_BLOCK + tid];
    }
    r.PAR_field[gidx] = PARv;
    if (r.T_field != nullptr) r.T_field[gidx] = Tv;
}

// tracer order = required_biogeochemical_tracers (NutrientsPlanktonDetritus.jl:69-74)
enum Role { R_NO3, R_NH4, R_FE, R_N, R_P, R_Z, R_T, R_D, R_SPOM, R_BPOM, R_DOM, R_SPOC, R_BPOC, R_DOC, R_DIC, R_ALK, R_O2 };

static int npd_layout(const obm_npd_params* p, int* roles, char (*names)[16]) {
    int n = 0;
    auto add = [&](int role, const char* nm) {
        if (roles) roles[n] = role;
        if (names) {
            memset(names[n], 0, 16);
            memcpy(names[n], nm, strnlen(nm, 15));  // ≤ 15 bytes of UTF-8, NUL-terminated by the memset
        }
        n++;
    };
    switch (p->nutrients) {
        case OBM_NUT_NUTRIENT: add(R_N, "N"); break;
        case OBM_NUT_NITRATE_AMMONIA: add(R_NO3, "NO₃"); add(R_NH4, "NH₄"); break;
        case OBM_NUT_NITRATE_AMMONIA_IRON: add(R_NO3, "NO₃"); add(R_NH4, "NH₄"); add(R_FE, "Fe"); break;
        default: set_error("unknown nutrients enum %d", p->nutrients); return OBM_EENUM;
    }
    add(R_P, "P");
    add(R_Z, "Z");
    if (p->has_temperature_coefficient) add(R_T, "T");
    switch (p->detritus) {
        case OBM_DET_NONE: break;
        case OBM_DET_DETRITUS: add(R_D, "D"); break;
        case OBM_DET_TWO_PARTICLE: add(R_SPOM, "sPOM"); add(R_BPOM, "bPOM"); add(R_DOM, "DOM"); break;
        case OBM_DET_VARIABLE_REDFIELD:
            add(R_SPOC, "sPOC"); add(R_BPOC, "bPOC"); add(R_DOC, "DOC");
            add(R_SPOM, "sPON"); add(R_BPOM, "bPON"); add(R_DOM, "DON");
            break;
        default: set_error("unknown detritus enum %d", p->detritus); return OBM_EENUM;
    }
    const int N = p->carbonate_replicates;
    if (N < 0 || N > MAX_REPLICATES) {
        set_error("carbonate_replicates = %d outside [0, %d]", N, MAX_REPLICATES);
        return N < 0 ? OBM_ESIZE : OBM_ENOTIMPL;
    }
    if (N == 1) {
        add(R_DIC, "DIC");
        add(R_ALK, "Alk");
    } else if (N > 1) {
        char buf[16];
        for (int r = 1; r <= N; r++) { snprintf(buf, 16, "DIC%d", r); add(R_DIC, buf); }
        for (int r = 1; r <= N; r++) { snprintf(buf, 16, "Alk%d", r); add(R_ALK, buf); }
    }
    if (p->oxygen) add(R_O2, "O₂");
    return n;
}

template <int NUT, bool ENSEMBLE>
static void launch_det(const NpdArgs& a, int det, dim3 blocks, cudaStream_t s) {
    switch (det) {
        case OBM_DET_NONE: npd_tendency_kernel<NUT, OBM_DET_NONE, ENSEMBLE><<<blocks, 256, 0, s>>>(a); break;
        case OBM_DET_DETRITUS: npd_tendency_kernel<NUT, OBM_DET_DETRITUS, ENSEMBLE><<<blocks, 256, 0, s>>>(a); break;
        case OBM_DET_TWO_PARTICLE: npd_tendency_kernel<NUT, OBM_DET_TWO_PARTICLE, ENSEMBLE><<<blocks, 256, 0, s>>>(a); break;
        default: npd_tendency_kernel<NUT, OBM_DET_VARIABLE_REDFIELD, ENSEMBLE><<<blocks, 256, 0, s>>>(a); break;
    }
}
template <bool ENSEMBLE>
static void launch_nut(const NpdArgs& a, int nut, int det, dim3 blocks, cudaStream_t s) {
    switch (nut) {
        case OBM_NUT_NUTRIENT: launch_det<OBM_NUT_NUTRIENT, ENSEMBLE>(a, det, blocks, s); break;
        case OBM_NUT_NITRATE_AMMONIA: launch_det<OBM_NUT_NITRATE_AMMONIA, ENSEMBLE>(a, det, blocks, s); break;
        default: launch_det<OBM_NUT_NITRATE_AMMONIA_IRON, ENSEMBLE>(a, det, blocks, s); break;
    }
}

template <int NUT, bool ENSEMBLE>
static void run_det(const NpdArgs& a, const RunArgs& r, int det, unsigned blocks, size_t smem, cudaStream_t s) {
    switch (det) {
        case OBM_DET_NONE: npd_box_run_kernel<NUT, OBM_DET_NONE, ENSEMBLE><<<blocks, RUN_BLOCK, smem, s>>>(a, r); break;
        case OBM_DET_DETRITUS: npd_box_run_kernel<NUT, OBM_DET_DETRITUS, ENSEMBLE><<<blocks, RUN_BLOCK, smem, s>>>(a, r); break;
        case OBM_DET_TWO_PARTICLE: npd_box_run_kernel<NUT, OBM_DET_TWO_PARTICLE, ENSEMBLE><<<blocks, RUN_BLOCK, smem, s>>>(a, r); break;
        default: npd_box_run_kernel<NUT, OBM_DET_VARIABLE_REDFIELD, ENSEMBLE><<<blocks, RUN_BLOCK, smem, s>>>(a, r); break;
    }
}
template <bool ENSEMBLE>
static void run_nut(const NpdArgs& a, const RunArgs& r, int nut, int det, unsigned blocks, size_t smem, cudaStream_t s) {
    switch (nut) {
        case OBM_NUT_NUTRIENT: run_det<OBM_NUT_NUTRIENT, ENSEMBLE>(a, r, det, blocks, smem, s); break;
        case OBM_NUT_NITRATE_AMMONIA: run_det<OBM_NUT_NITRATE_AMMONIA, ENSEMBLE>(a, r, det, blocks, smem, s); break;
        default: run_det<OBM_NUT_NITRATE_AMMONIA_IRON, ENSEMBLE>(a, r, det, blocks, smem, s); break;
    }
}

}  // namespace obm

using namespace obm;

extern "C" int obm_npd_tracer_names(const obm_npd_params* p, char (*names)[16]) {
    OBM_REQUIRE(p != nullptr, OBM_ENULL, "params is NULL");
    return npd_layout(p, nullptr, names);
}

extern "C" int obm_npd_param_index(const char* name) {
    OBM_REQUIRE(name != nullptr, OBM_ENULL, "obm_npd_param_index: name is NULL");
#define X(n) if (strcmp(name, #n) == 0) return NPD_PARAM_##n;
    OBM_NPD_DOUBLE_PARAMS(X)
#undef X
    set_error("obm_npd_param_index: obm_npd_params has no double member '%s'", name);
    return OBM_EENUM;
}

struct StepSpec {  // the fused tendency + substep launch; null: tendencies only
    double* const* U;
    double* const* Gm;
    double dt, gamma, zeta;
    int has_zeta, store_gn;
};
struct RunSpec {  // obm_npd_box_run: the whole run in one launch
    RunArgs r;
    double* const* snapshots;  // per tracer (tracer order), nullable entries / nullable table
};
static int npd_launch(const obm_grid* grid, const obm_npd_params* p, int nvary, const int32_t* which, const double* values,
                      const double* const* tracers, const double* PAR, double* const* G, int accumulate, void* stream,
                      bool ensemble, const StepSpec* step = nullptr, RunSpec* run = nullptr) {
    OBM_REQUIRE(p != nullptr && tracers != nullptr && (G != nullptr || step != nullptr) && PAR != nullptr, OBM_ENULL,
                "obm_npd_tendencies: params / tracers / G / PAR is NULL");
    OBM_REQUIRE(step == nullptr || (step->U != nullptr && step->Gm != nullptr), OBM_ENULL,
                "obm_npd_tendencies_substep: tracers / G⁻ table is NULL");
    OBM_REQUIRE(step == nullptr || G != nullptr || (!accumulate && !step->store_gn), OBM_ENULL,
                "obm_npd_tendencies_substep: accumulate / store_Gn need the Gⁿ table");
    NpdArgs a;
    memset(&a, 0, sizeof(a));
    int rc = make_dims(grid, &a.d, false);
    if (rc) return rc;
    if (ensemble) {
        OBM_REQUIRE(nvary >= 0 && nvary <= OBM_NPD_MAX_VARIED, OBM_ESIZE, "obm_npd_tendencies_ensemble: nvary = %d outside [0, %d]",
                    nvary, OBM_NPD_MAX_VARIED);
        OBM_REQUIRE(nvary == 0 || (which != nullptr && values != nullptr), OBM_ENULL,
                    "obm_npd_tendencies_ensemble: which / values is NULL");
        for (int v = 0; v < nvary; v++) {
            OBM_REQUIRE(which[v] >= 0 && which[v] < NPD_PARAM_COUNT, OBM_EENUM,
                        "obm_npd_tendencies_ensemble: which[%d] = %d is not a parameter index (0 … %d)", v, which[v],
                        NPD_PARAM_COUNT - 1);
            // temperature_coefficient changes the tracer list through has_temperature_coefficient only; its value may vary
            a.which[v] = which[v];
        }
        a.nvary = nvary;
        a.values = values;
    }
    int roles[OBM_NPD_MAX_TRACERS];
    const int nt = npd_layout(p, roles, nullptr);
    if (nt < 0) return nt;
    a.p = *p;
    a.PAR = PAR;
    a.accumulate = accumulate ? 1 : 0;
    if (step) {
        a.step.on = 1;
        a.step.has_zeta = step->has_zeta ? 1 : 0;
        a.step.store_gn = step->store_gn ? 1 : 0;
        a.step.dt = step->dt; a.step.gamma = step->gamma; a.step.zeta = step->zeta;
    }
    int nd = 0, na = 0;
    for (int n = 0; n < nt; n++) {
        const double* c = tracers[n];
        double* g = G ? G[n] : nullptr;
        double* u = step ? step->U[n] : nullptr;          // stepped only where BOTH the tracer and its G⁻ are given
        double* m = step ? step->Gm[n] : nullptr;
        if (step && (u == nullptr || (m == nullptr))) u = m = nullptr;
        OBM_REQUIRE(!step || u == nullptr || u == c, OBM_ESIZE,
                    "obm_npd_tendencies_substep: tracers[%d] and the stepped field differ", n);
        const bool read = !(roles[n] == R_DIC || roles[n] == R_ALK || roles[n] == R_O2);  // one-way coupled: never read
        OBM_REQUIRE(!read || c != nullptr, OBM_ENULL, "obm_npd_tendencies: tracers[%d] is NULL", n);
        switch (roles[n]) {
            case R_NO3: a.NO3 = c; a.gNO3 = g; a.uNO3 = u; a.mNO3 = m; break;
            case R_NH4: a.NH4 = c; a.gNH4 = g; a.uNH4 = u; a.mNH4 = m; break;
            case R_FE: a.Fe = c; a.gFe = g; a.uFe = u; a.mFe = m; break;
            case R_N: a.N = c; a.gN = g; a.uN = u; a.mN = m; break;
            case R_P: a.P = c; a.gP = g; a.uP = u; a.mP = m; break;
            case R_Z: a.Z = c; a.gZ = g; a.uZ = u; a.mZ = m; break;
            case R_T: a.T = c; break;  // no biogeochemical tendency: zero(grid) NutrientsPlanktonDetritus.jl:88
            case R_D: a.D = c; a.gD = g; a.uD = u; a.mD = m; break;
            case R_SPOM: a.sPOM = c; a.gsPOM = g; a.usPOM = u; a.msPOM = m; break;
            case R_BPOM: a.bPOM = c; a.gbPOM = g; a.ubPOM = u; a.mbPOM = m; break;
            case R_DOM: a.DOM = c; a.gDOM = g; a.uDOM = u; a.mDOM = m; break;
            case R_SPOC: a.sPOC = c; a.gsPOC = g; a.usPOC = u; a.msPOC = m; break;
            case R_BPOC: a.bPOC = c; a.gbPOC = g; a.ubPOC = u; a.mbPOC = m; break;
            case R_DOC: a.DOC = c; a.gDOC = g; a.uDOC = u; a.mDOC = m; break;
            case R_DIC: a.gDIC[nd] = g; a.uDIC[nd] = u; a.mDIC[nd] = m; nd++; break;
            case R_ALK: a.gAlk[na] = g; a.uAlk[na] = u; a.mAlk[na] = m; na++; break;
            case R_O2: a.gO2 = g; a.uO2 = u; a.mO2 = m; break;
        }
    }
    a.nrep = nd;
    cudaStream_t s = (cudaStream_t)stream;
    if (run != nullptr) {
        RunArgs& r = run->r;
        OBM_REQUIRE(a.d.Ny == 1 && a.d.Nz == 1, OBM_ENOTIMPL, "obm_npd_box_run: boxes lie along x (Ny = Nz = 1), got %d x %d x %d",
                    a.d.Nx, a.d.Ny, a.d.Nz);
        r.ncells = a.d.i1 - a.d.i0;
        // slot order of the kernel's staging (OBM_STAGE there): every tracer the cell READS must be stepped — a prescribed
        // P or Z would have to come from a table this entry point does not take
        double* slot[RUN_MAX_SLOTS];
        int ns = 0;
#define OBM_SLOT(name)                                                                                                  \
    if (a.u##name != nullptr) slot[ns++] = a.u##name;                                                                     \
    else OBM_REQUIRE(a.name == nullptr, OBM_ENOTIMPL, "obm_npd_box_run: tracer " #name " is read but not stepped (prescribed?)");
        OBM_SLOT(NO3) OBM_SLOT(NH4) OBM_SLOT(Fe) OBM_SLOT(N) OBM_SLOT(P) OBM_SLOT(Z) OBM_SLOT(D) OBM_SLOT(sPOM)
        OBM_SLOT(bPOM) OBM_SLOT(DOM) OBM_SLOT(sPOC) OBM_SLOT(bPOC) OBM_SLOT(DOC)
#undef OBM_SLOT
        for (int q = 0; q < a.nrep; q++) {
            OBM_REQUIRE(a.uDIC[q] && a.uAlk[q], OBM_ENOTIMPL, "obm_npd_box_run: DIC / Alk replicate %d is not stepped", q);
            slot[ns++] = a.uDIC[q];
            slot[ns++] = a.uAlk[q];
        }
        if (a.uO2 != nullptr) slot[ns++] = a.uO2;
        r.nslots = ns;
        for (int q = 0; q < ns; q++) {
            r.snap[q] = nullptr;
            for (int n = 0; n < nt && run->snapshots != nullptr; n++)
                if (step->U[n] == slot[q]) r.snap[q] = run->snapshots[n];
        }
        r.T_field = const_cast<double*>(a.T);
        OBM_REQUIRE(r.T_table == nullptr || r.T_field != nullptr, OBM_ESIZE,
                    "obm_npd_box_run: a temperature table for a model without temperature dependence");
        const unsigned blocks = (unsigned)((r.ncells + RUN_BLOCK - 1) / RUN_BLOCK);
        const size_t smem = ((size_t)2 * ns + 2) * RUN_BLOCK * sizeof(double);
        if (ensemble) run_nut<true>(a, r, p->nutrients, p->detritus, blocks, smem, s);
        else run_nut<false>(a, r, p->nutrients, p->detritus, blocks, smem, s);
        return launch_status("npd_box_run_kernel");
    }
    const dim3 blocks = cell_grid(a.d, 256);
    if (ensemble) launch_nut<true>(a, p->nutrients, p->detritus, blocks, s);
    else launch_nut<false>(a, p->nutrients, p->detritus, blocks, s);
    return launch_status("npd_tendency_kernel");
}

extern "C" int obm_npd_tendencies(const obm_grid* grid, const obm_npd_params* p, const double* const* tracers,
                                  const double* PAR, double* const* G, int accumulate, void* stream) {
    return npd_launch(grid, p, 0, nullptr, nullptr, tracers, PAR, G, accumulate, stream, false);
}

extern "C" int obm_npd_tendencies_ensemble(const obm_grid* grid, const obm_npd_params* p, int nvary, const int32_t* which,
                                           const double* values, const double* const* tracers, const double* PAR,
                                           double* const* G, int accumulate, void* stream) {
    return npd_launch(grid, p, nvary, which, values, tracers, PAR, G, accumulate, stream, true);
}

// f-2: the tendencies and the tracer update in ONE launch (see include/obm_b200.h)
extern "C" int obm_npd_tendencies_substep(const obm_grid* grid, const obm_npd_params* p, int nvary, const int32_t* which,
                                          const double* values, double* const* tracers, const double* PAR, double* const* G,
                                          int accumulate, int store_Gn, double* const* Gm, double dt, double gamma, double zeta,
                                          int has_zeta, void* stream) {
    const StepSpec step{tracers, Gm, dt, gamma, zeta, has_zeta, store_Gn};
    return npd_launch(grid, p, nvary, which, values, tracers, PAR, G, accumulate, stream, nvary > 0, &step);
}

// f-3: every stage of every time step of a box-model ensemble in ONE launch (see include/obm_b200.h)
extern "C" int obm_npd_box_run(const obm_grid* grid, const obm_npd_params* p, int nvary, const int32_t* which, const double* values,
                               double* const* tracers, double* const* Gm, double* PAR, const double* PAR_table, int PAR_per_box,
                               const double* T_table, int T_per_box, int nsteps, int nstages, const double* gamma,
                               const double* zeta, double dt, int output_every, double* const* snapshots, void* stream) {
    OBM_REQUIRE(PAR && PAR_table && gamma && zeta, OBM_ENULL, "obm_npd_box_run: PAR / PAR_table / gamma / zeta is NULL");
    OBM_REQUIRE(nsteps >= 0 && nstages >= 1 && nstages <= 3 && output_every >= 0, OBM_ESIZE,
                "obm_npd_box_run: nsteps = %d, nstages = %d, output_every = %d", nsteps, nstages, output_every);
    if (nsteps == 0) return 0;
    RunSpec run;
    memset(&run, 0, sizeof(run));
    RunArgs& r = run.r;
    r.nsteps = nsteps; r.nstages = nstages; r.output_every = output_every;
    for (int q = 0; q < nstages; q++) {
        r.gamma[q] = gamma[q];
        r.has_zeta[q] = zeta[q] == zeta[q] ? 1 : 0;  // NaN: a stage without a ζ term (the first of RK3, forward Euler)
        r.zeta[q] = r.has_zeta[q] ? zeta[q] : 0.0;
    }
    r.PAR_field = PAR; r.PAR_table = PAR_table; r.PAR_per_box = PAR_per_box ? 1 : 0;
    r.T_table = T_table; r.T_per_box = T_per_box ? 1 : 0;
    run.snapshots = snapshots;
    const StepSpec step{tracers, Gm, dt, gamma[0], 0.0, 0, 0};
    return npd_launch(grid, p, nvary, which, values, tracers, PAR, nullptr, 0, stream, nvary > 0, &step, &run);
}
