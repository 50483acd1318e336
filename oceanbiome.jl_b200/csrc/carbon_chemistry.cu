// carbon_chemistry.cu — kernels around carbon_chemistry.cuh: flat sweep (C5 of BASELINE.json,
// validation/carbon_chemistry.jl style) and the gridded calcite-saturation state of PISCES
// (PISCES/compute_calcite_saturation.jl:9-37).  One thread per cell; FP64-pipe bound.
#include "carbon_chemistry.cuh"

namespace obm {

struct SweepArgs {
    long long n;
    const double *T, *S, *DIC, *Alk, *P, *sil, *phos, *pH;
    double* out;
    int output_kind, iterations;
    double H_init;  // 10^(−initial_pH_guess), host-evaluated
};

template <bool HAS_P>
__global__ void __launch_bounds__(128) carbon_sweep_kernel(const __grid_constant__ SweepArgs a) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n) return;
    const double P = HAS_P ? a.P[c] : 0.0;
    a.out[c] = cc::solve<HAS_P>(a.output_kind, a.T[c], a.S[c], a.DIC[c], a.Alk ? a.Alk[c] : 0.0, P, a.sil != nullptr,
                                a.sil ? a.sil[c] : 0.0, a.phos != nullptr, a.phos ? a.phos[c] : 0.0, a.pH != nullptr,
                                a.pH ? a.pH[c] : 0.0, a.H_init, a.iterations);
}

struct OmegaArgs {
    GridDims d;
    const double *T, *S, *DIC, *Alk, *Si;
    double *Omega, *Hst;
    int iterations;
    double H_init;  // 10^(−initial_pH_guess), host-evaluated
};

__global__ void __launch_bounds__(128) calcite_saturation_kernel(const __grid_constant__ OmegaArgs a) {
#if OBM_CC_LEVEL
    __shared__ cc::LevelTables level;  // the block's z-level fixes the pressure (see carbon_chemistry.cuh)
    cc::fill_level_entry(level, fabs(a.d.zc[blockIdx.z]) * 9.80665 * 1026.0 / 100000.0, threadIdx.x);
    __syncthreads();
    const cc::LevelTables* lvl = &level;
#else
    const cc::LevelTables* lvl = nullptr;
#endif
    int i, j, k;
    if (!thread_cell(a.d, i, j, k)) return;
    const long long idx = cell_index(a.d, i, j, k);
    // P = abs(z) * g * 1026 / 100000 with g = Oceananigans.defaults.gravitational_acceleration
    const double P = fabs(a.d.zc[k]) * 9.80665 * 1026.0 / 100000.0;
    a.Omega[idx] = cc::solve<true>(OBM_CC_OMEGA_CALCITE, a.T[idx], a.S[idx], a.DIC[idx], a.Alk[idx], P, true, a.Si[idx],
                                   false, 0.0, false, 0.0, a.H_init, a.iterations, a.Hst ? a.Hst + idx : nullptr, lvl);
}

static void defaults(const obm_carbchem_params* p, int* iterations, double* pH0) {
    *iterations = (p && p->newton_iterations > 0) ? p->newton_iterations : 12;
    *pH0 = pow(10.0, -((p && p->initial_pH_guess > 0) ? p->initial_pH_guess : 8.0));
}

}  // namespace obm

using namespace obm;

extern "C" int obm_carbon_chemistry(int64_t n, const obm_carbchem_params* p, const double* T, const double* S,
                                    const double* DIC, const double* Alk, const double* P_bar, const double* silicate,
                                    const double* phosphate, const double* pH, int output_kind, double* out,
                                    void* stream) {
    OBM_REQUIRE(n >= 0, OBM_ESIZE, "obm_carbon_chemistry: n = %lld", (long long)n);
    if (n == 0) return 0;
    OBM_REQUIRE(T && S && DIC && out, OBM_ENULL, "obm_carbon_chemistry: T / S / DIC / out is NULL");
    OBM_REQUIRE(Alk || pH, OBM_ENULL, "obm_carbon_chemistry: one of Alk, pH must be given");
    OBM_REQUIRE(output_kind >= OBM_CC_FCO2 && output_kind <= OBM_CC_OMEGA_CALCITE, OBM_EENUM,
                "obm_carbon_chemistry: unknown output kind %d", output_kind);
    SweepArgs a;
    a.n = n; a.T = T; a.S = S; a.DIC = DIC; a.Alk = Alk; a.P = P_bar; a.sil = silicate; a.phos = phosphate; a.pH = pH;
    a.out = out;
    a.output_kind = output_kind;
    defaults(p, &a.iterations, &a.H_init);
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (P_bar) carbon_sweep_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(a);
    else carbon_sweep_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("carbon_sweep_kernel");
}

extern "C" int obm_calcite_saturation(const obm_grid* grid, const obm_carbchem_params* p, const double* T,
                                      const double* S, const double* DIC, const double* Alk, const double* Si,
                                      double* Omega, double* H_state, void* stream) {
    OBM_REQUIRE(T && S && DIC && Alk && Si && Omega, OBM_ENULL, "obm_calcite_saturation: a field pointer is NULL");
    OmegaArgs a;
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    a.T = T; a.S = S; a.DIC = DIC; a.Alk = Alk; a.Si = Si; a.Omega = Omega; a.Hst = H_state;
    defaults(p, &a.iterations, &a.H_init);
    calcite_saturation_kernel<<<cell_grid(a.d, 128), 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("calcite_saturation_kernel");
}
