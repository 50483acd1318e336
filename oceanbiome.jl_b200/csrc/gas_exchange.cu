// gas_exchange.cu — air–sea gas-exchange flux for every surface column in one x–y launch
// (src/Models/GasExchange/gas_exchange.jl:26-38).  In the reference this runs inside Oceananigans'
// boundary-condition kernel, one Newton solve per surface cell per tracer with a flux BC; here it is
// one thread per column writing the flux plane (and, optionally, the top-cell tendency in place).
// FP64-pipe bound (the carbonate solve); Nx·Ny threads, so the launch is a small fraction of a stage.
#include "carbon_chemistry.cuh"

namespace obm {

struct GasArgs {
    GridDims d;
    obm_gas_exchange_params p;
    const double *T, *S, *tracer, *DIC, *Alk, *sil, *phos, *wind, *air;
    double *flux, *G;
    int iterations;
    double H_init;  // 10^(−initial_pH_guess), host-evaluated
};

// Base.Math.pow_body(x, 4): compensated squaring (what Julia's `x^4` evaluates for Float64)
__device__ __forceinline__ double jl_pow4(double x) {
    const double x2 = x * x, l2 = fma(x, x, -x2);
    const double x4 = x2 * x2;
    const double l4 = fma(x2, x2, -x4) + x2 * 2 * l2;
    return (isfinite(x4) && isfinite(l4)) ? x4 + l4 : x4;
}

// PolynomialParameterisation{N} — generic_parameterisations.jl:24-36 (left-to-right sums, no Horner)
__device__ __forceinline__ double polynomial(int order, const double* c, double x) {
    double y = c[0];
    if (order >= 1) y = __dadd_rn(y, __dmul_rn(c[1], x));
    if (order >= 2) y = __dadd_rn(y, __dmul_rn(c[2], __dmul_rn(x, x)));
    if (order >= 3) y = __dadd_rn(y, __dmul_rn(c[3], __dmul_rn(__dmul_rn(x, x), x)));
    if (order >= 4) y = __dadd_rn(y, __dmul_rn(c[4], jl_pow4(x)));
    return y;
}

__global__ void __launch_bounds__(128) gas_exchange_kernel(const __grid_constant__ GasArgs a) {
    const int ii = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (ii >= a.d.i1 - a.d.i0) return;
    const int i = a.d.i0 + ii, j = a.d.j0 + (int)blockIdx.y;
    const long long idx = cell_index(a.d, i, j, a.d.Nz - 1);
    const long long pidx = plane_index(a.d, i, j);
    const obm_gas_exchange_params& p = a.p;

    const double T = a.T[idx], S = a.S[idx];
    const double u10 = a.wind ? a.wind[pidx] : p.wind_speed;
    double air = a.air ? a.air[pidx] : p.air_concentration;

    // k = k₆₆₀(u₁₀) / √(Sc(T)/660) · solubility(T, S) — gas_transfer_velocity.jl:32-33
    double k = polynomial(p.k660_order, p.k660, u10) / sqrt(polynomial(4, p.schmidt, T) / 660.0);
    if (p.solubility_kind == OBM_GE_SOLUBILITY_K0_RHO) {  // gas_solubility.jl:65 (density at Pbar = 0)
        const double Tk = T + 273.15;
        k = k * (cc::K0(Tk, log(Tk), S) * cc::teos10_rho(T, S, 0.0) / 1000.0);
    } else {
        k = k * 1.0;
    }

    if (p.air_kind == OBM_GE_AIR_WANNINKHOF92) {  // gas_solubility.jl:24-25, :34-47 (B2 used twice, as found)
        const double Tk = T + 273.15, Tk_100 = Tk / 100.0;
        const double beta = exp(p.w92[0] + p.w92[1] / Tk_100 + p.w92[2] * log(Tk_100) +
                                S * (p.w92[3] + p.w92[4] * Tk_100 + p.w92[4] * (Tk_100 * Tk_100)));
        air = air * (beta / Tk);
    }

    double water;
    if (p.water_kind == OBM_GE_WATER_PCO2) {  // carbon_dioxide_concentration.jl:48-60
        const bool sp = p.use_silicate_phosphate != 0;
        const double sil = sp ? (a.sil ? a.sil[idx] : p.silicate) : 0.0;
        const double phos = sp ? (a.phos ? a.phos[idx] : p.phosphate) : 0.0;
        water = cc::solve<false>(OBM_CC_PCO2, T, S, a.DIC[idx], a.Alk[idx], 0.0, sp, sil, sp, phos, false, 0.0,
                                 a.H_init, a.iterations);
    } else {
        water = a.tracer[idx];
    }

    const double flux = k * (water - air);
    if (a.flux) a.flux[pidx] = flux;
    if (a.G) a.G[idx] -= flux / (a.d.zf[a.d.Nz] - a.d.zf[a.d.Nz - 1]);  // Δzᶜ(Nz)
}

}  // namespace obm

using namespace obm;

extern "C" int obm_gas_exchange_flux(const obm_grid* grid, const obm_gas_exchange_params* p, const double* T,
                                     const double* S, const double* tracer, const double* DIC, const double* Alk,
                                     const double* silicate_f, const double* phosphate_f, const double* wind_speed_xy,
                                     const double* air_concentration_xy, double* flux_xy, double* G_top, void* stream) {
    OBM_REQUIRE(p != nullptr, OBM_ENULL, "obm_gas_exchange_flux: params is NULL");
    OBM_REQUIRE(T && S, OBM_ENULL, "obm_gas_exchange_flux: T / S is NULL");
    OBM_REQUIRE(flux_xy || G_top, OBM_ENULL, "obm_gas_exchange_flux: one of flux_xy, G_top must be given");
    OBM_REQUIRE(p->water_kind == OBM_GE_WATER_TRACER || p->water_kind == OBM_GE_WATER_PCO2, OBM_EENUM,
                "obm_gas_exchange_flux: unknown water_kind %d", p->water_kind);
    OBM_REQUIRE(p->air_kind == OBM_GE_AIR_PLAIN || p->air_kind == OBM_GE_AIR_WANNINKHOF92, OBM_EENUM,
                "obm_gas_exchange_flux: unknown air_kind %d", p->air_kind);
    OBM_REQUIRE(p->solubility_kind == OBM_GE_SOLUBILITY_ONE || p->solubility_kind == OBM_GE_SOLUBILITY_K0_RHO, OBM_EENUM,
                "obm_gas_exchange_flux: unknown solubility_kind %d", p->solubility_kind);
    OBM_REQUIRE(p->k660_order >= 0 && p->k660_order <= 3, OBM_ESIZE, "obm_gas_exchange_flux: k660_order %d not in 0…3",
                p->k660_order);
    if (p->water_kind == OBM_GE_WATER_PCO2)
        OBM_REQUIRE(DIC && Alk, OBM_ENULL, "obm_gas_exchange_flux: DIC / Alk is NULL for a pCO2 water value");
    else
        OBM_REQUIRE(tracer, OBM_ENULL, "obm_gas_exchange_flux: tracer is NULL");
    GasArgs a;
    int rc = make_dims(grid, &a.d, G_top != nullptr);
    if (rc) return rc;
    a.p = *p;
    a.T = T; a.S = S; a.tracer = tracer; a.DIC = DIC; a.Alk = Alk; a.sil = silicate_f; a.phos = phosphate_f;
    a.wind = wind_speed_xy; a.air = air_concentration_xy; a.flux = flux_xy; a.G = G_top;
    a.iterations = p->carbon_chemistry.newton_iterations > 0 ? p->carbon_chemistry.newton_iterations : 12;
    a.H_init = pow(10.0, -(p->carbon_chemistry.initial_pH_guess > 0 ? p->carbon_chemistry.initial_pH_guess : 8.0));
    const unsigned chunks = (unsigned)((a.d.i1 - a.d.i0 + 127) / 128);
    const unsigned ny = (unsigned)(a.d.j1 - a.d.j0);
    OBM_REQUIRE(ny <= 65535u, OBM_ESIZE, "obm_gas_exchange_flux: more than 65535 rows per launch (%u); restrict j", ny);
    gas_exchange_kernel<<<dim3(chunks, ny, 1), 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("gas_exchange_kernel");
}
