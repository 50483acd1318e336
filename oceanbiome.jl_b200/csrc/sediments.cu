// sediments.cu — bottom-sediment coupling as two fused :xy kernels (one per hook).
//
// Replaces the ≈ 15 tiny launches of src/Sediments/ per stage (K7–K11 of SURVEY §2.2: one gather per
// tracked tracer, one flux kernel per sinking tracer, one tendency kernel and one stepping kernel per
// sediment pool, one coupling kernel per coupled tracer) with:
//   sediment_state_kernel       gather + sinking fluxes + pool step (AB2 / RK3) + tendency cache + new Gⁿ
//   sediment_tendency_kernel    G[i, j, k_bottom] += flux / Δz for every coupled tracer
// One thread per column, x fastest (the bottom plane of a flat-bottomed grid is contiguous).
// Equations: src/Models/Sediments/instant_remineralisation.jl:103-125, simple_multi_G.jl:165-427.
// Assumptions on the Oceananigans side (flux operator, stepping order) are listed in DESIGN.md §sediments;
// the reference leaves this path unpinned (its testset is commented out).
#include <string.h>

#include "obm_common.cuh"

namespace obm {

constexpr double DAY = 86400.0;

struct SedArgs {
    GridDims d;
    obm_sediment_params p;
    obm_sediment_fields f;
    double dt, chi;
    int np, nc, nt;  // pools, coupled tracers, tracked tracers
};

struct SedPoint {
    double pool[OBM_SED_MAX_POOLS];
    double NO3, NH4, O2, fN, fC;
};

__device__ __forceinline__ double sinking_flux(const SedArgs& a, const double* C, const double* w, long long idx) {
    const double wk = w[idx], CL = C[idx - a.d.sz], CR = C[idx];
    if (a.p.advection == OBM_ADV_UPWIND1) return -(((wk + fabs(wk)) * CL + (wk - fabs(wk)) * CR) / 2);
    return -(wk * ((CL + CR) / 2));
}

__device__ __forceinline__ int bottom_k(const SedArgs& a, long long pl) {
    return a.f.bottom_indices_xy ? (int)(a.f.bottom_indices_xy[pl] - 1) : 0;
}

__device__ __forceinline__ void gather(const SedArgs& a, int i, int j, long long pl, int k, SedPoint& c, double* fluxes) {
    const long long idx = cell_index(a.d, i, j, k);
#pragma unroll
    for (int n = 0; n < OBM_SED_MAX_POOLS; n++) c.pool[n] = n < a.np ? a.f.pools[n][pl] : 0.0;
    c.NO3 = c.NH4 = c.O2 = 0.0;
    if (a.nt) { c.NO3 = a.f.NO3[idx]; c.NH4 = a.f.NH4[idx]; c.O2 = a.f.O2[idx]; }
    c.fN = c.fC = 0.0;
#pragma unroll
    for (int n = 0; n < OBM_SED_MAX_SINKING; n++)
        if (n < a.p.nsinking_nitrogen) {
            const double fl = sinking_flux(a, a.f.sinking[n], a.f.sinking_w[n], idx);
            if (fluxes) fluxes[n] = fl;
            c.fN = n == 0 ? fl : c.fN + fl;
        }
#pragma unroll
    for (int n = 0; n < OBM_SED_MAX_SINKING; n++)
        if (n < a.p.nsinking_carbon) {
            const int q = a.p.nsinking_nitrogen + n;
            const double fl = sinking_flux(a, a.f.sinking[q], a.f.sinking_w[q], idx);
            if (fluxes) fluxes[OBM_SED_MAX_SINKING + n] = fl;
            c.fC = n == 0 ? fl : c.fC + fl;
        }
}

__device__ __forceinline__ double burial_efficiency(const obm_sediment_params& s, double flux) {
    const double q = flux / (s.burial_efficiency_half_saturation + flux);
    return s.burial_efficiency_constant1 + s.burial_efficiency_constant2 * (q * q);
}

__device__ __forceinline__ double pool_tendency(const obm_sediment_params& s, const SedPoint& c, int n) {
    if (s.model == OBM_SED_INSTANT_REMINERALISATION) return burial_efficiency(s, c.fN) * c.fN;
    const double fr = s.refactory_fraction;
    switch (n) {
        case 0: return (1 - fr) * s.slow_fraction * c.fN - s.slow_decay_rate * c.pool[0];
        case 1: return (1 - fr) * s.fast_fraction * c.fN - s.fast_decay_rate * c.pool[1];
        case 2: return fr * c.fN;
        case 3: return (1 - fr) * s.slow_fraction * c.fC - s.slow_decay_rate * c.pool[3];
        case 4: return (1 - fr) * s.fast_fraction * c.fC - s.fast_decay_rate * c.pool[4];
        default: return fr * c.fC;
    }
}

// the Soetaert (2000) meta-model fractions share log(Cr·day), log(k·day), log k, log O₂, log NO₃, log NH₄
__device__ __forceinline__ void coupled_fluxes(const obm_sediment_params& s, const SedPoint& c, double* out) {
    if (s.model == OBM_SED_INSTANT_REMINERALISATION) {
        out[0] = (1 - burial_efficiency(s, c.fN)) * c.fN;
        return;
    }
    const double Ns = c.pool[0], Nf = c.pool[1];
    const double Nr = s.slow_decay_rate * Ns + s.fast_decay_rate * Nf;
    double Cr, k;
    if (s.carbon) {
        const double Cs = c.pool[3], Cf = c.pool[4];
        Cr = s.slow_decay_rate * Cs + s.fast_decay_rate * Cf;
        k = Cr / (Cs + Cf + eps0());
    } else {
        const double R = s.sinking_redfield;
        Cr = Nr * R;
        const double Cs = Ns * R, Cf = Nf * R;
        k = (s.slow_decay_rate * Cs + s.fast_decay_rate * Cf) / (Cs + Cf + eps0());
    }
    const double kO2 = s.anoxia_half_saturation;
    const double lC = log(Cr * DAY), lkd = log(k * DAY), lk = log(k), lO = log(c.O2), lN3 = log(c.NO3), lN4 = log(c.NH4);
    const double oxf = c.O2 / (kO2 + c.O2);
    const double* q = s.nitrate_oxidation_params;
    double pn = exp(q[0] + q[1] * lC * lO + q[2] * (lC * lC) + q[3] * lkd * lN4 + q[4] * lC + q[5] * lC * lN4) / (Nr * DAY) * c.O2 / (kO2 + c.O2);
    pn = isfinite(pn) ? pn : 0.0;
    q = s.denitrification_params;
    double pnp = exp(q[0] + q[1] * lC + q[2] * (lN3 * lN3) + q[3] * (lC * lC) + q[4] * (lkd * lkd) + q[5] * lO * lk) / (Cr * DAY) * c.O2 / (kO2 + c.O2);
    pnp = isfinite(pnp) ? pnp : 0.0;
    q = s.anoxic_params;
    double pa = exp(q[0] + q[1] * lC + q[2] * (lC * lC) + q[3] * lkd + q[4] * lO * lk + q[5] * (lN3 * lN3)) / (Cr * DAY);
    pa = isfinite(pa) ? pa : 0.0;
    const double ps = 0.223 * pow(s.sedimentation_rate, 0.336);
    out[0] = pn * Nr - 0.8 * pnp * Cr;                                // NO₃
    out[1] = (1 - pn) * Nr + 0.8 * pnp * Cr;                          // NH₄
    out[2] = -(1 - pa * ps - pnp) * c.O2 / (kO2 + c.O2) * Cr - 2 * pn * Nr;  // O₂
    out[3] = Cr;                                                      // DIC
    (void)oxf;
}

__global__ void __launch_bounds__(128) sediment_state_kernel(const __grid_constant__ SedArgs a) {
    const int nx = a.d.i1 - a.d.i0;
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= (long long)nx * (a.d.j1 - a.d.j0)) return;
    const int jj = (int)(col / nx);
    const int i = a.d.i0 + (int)(col - (long long)jj * nx), j = a.d.j0 + jj;
    const long long pl = plane_index(a.d, i, j);
    const int k = bottom_k(a, pl);
    SedPoint c;
    double fluxes[2 * OBM_SED_MAX_SINKING];
    gather(a, i, j, pl, k, c, fluxes);  // K7, K8
    if (a.f.tracked_xy[0]) {
        int q = 0;
        if (a.nt) { a.f.tracked_xy[q++][pl] = c.NO3; a.f.tracked_xy[q++][pl] = c.NH4; a.f.tracked_xy[q++][pl] = c.O2; }
        for (int n = 0; n < a.p.nsinking_nitrogen; n++) a.f.tracked_xy[q++][pl] = fluxes[n];
        for (int n = 0; n < a.p.nsinking_carbon; n++) a.f.tracked_xy[q++][pl] = fluxes[OBM_SED_MAX_SINKING + n];
    }
    if (isfinite(a.dt) && a.p.timestepper == OBM_TS_RK3) {
        // time_step!(sediment_model::…{<:RungeKutta3TimeStepper}, Δt): the sediment takes a WHOLE three-stage step of
        // length Δt (the parent's last stage) inside this hook — per stage rk3_substep! (timesteppers.jl:46-73),
        // cache_previous_tendencies! (:84-96), update_state! → compute_sediment_tendencies! — with the tracked
        // tracers / fluxes held at the values gathered above.  The pools are per-column scalars, so all three stages
        // run in registers: one launch instead of 3 × (np step + np cache + np tendency) launches.
        const double gam[3] = {8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0}, zet[3] = {0.0, -17.0 / 60.0, -5.0 / 12.0};
        double Gn[OBM_SED_MAX_POOLS], Gm[OBM_SED_MAX_POOLS];
#pragma unroll
        for (int n = 0; n < OBM_SED_MAX_POOLS; n++)
            if (n < a.np) { Gn[n] = a.f.Gn[n][pl]; Gm[n] = a.f.Gm[n][pl]; }
#pragma unroll
        for (int st = 0; st < 3; st++) {
#pragma unroll
            for (int n = 0; n < OBM_SED_MAX_POOLS; n++)
                if (n < a.np) {
                    c.pool[n] += st == 0 ? a.dt * gam[0] * Gn[n] : a.dt * (gam[st] * Gn[n] + zet[st] * Gm[n]);
                    Gm[n] = Gn[n];
                }
#pragma unroll
            for (int n = 0; n < OBM_SED_MAX_POOLS; n++)
                if (n < a.np) Gn[n] = pool_tendency(a.p, c, n);
        }
#pragma unroll
        for (int n = 0; n < OBM_SED_MAX_POOLS; n++)
            if (n < a.np) { a.f.pools[n][pl] = c.pool[n]; a.f.Gm[n][pl] = Gm[n]; a.f.Gn[n][pl] = Gn[n]; }
        return;
    }
    if (isfinite(a.dt)) {  // time_step!(sediment_model::…{<:QuasiAdamsBashforth2TimeStepper}, Δt): K10 + tendency cache
#pragma unroll
        for (int n = 0; n < OBM_SED_MAX_POOLS; n++)
            if (n < a.np) {
                const double Gn = a.f.Gn[n][pl], Gm = a.f.Gm[n][pl];
                const double Gu = (1.5 + a.chi) * Gn - (a.chi != -0.5 ? (0.5 + a.chi) * Gm : 0.0);
                const double u = c.pool[n] + a.dt * Gu;
                c.pool[n] = u;
                a.f.pools[n][pl] = u;
                a.f.Gm[n][pl] = Gn;
            }
    }
#pragma unroll
    for (int n = 0; n < OBM_SED_MAX_POOLS; n++)
        if (n < a.np) a.f.Gn[n][pl] = pool_tendency(a.p, c, n);  // K9
}

__global__ void __launch_bounds__(128) sediment_tendency_kernel(const __grid_constant__ SedArgs a) {
    const int nx = a.d.i1 - a.d.i0;
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= (long long)nx * (a.d.j1 - a.d.j0)) return;
    const int jj = (int)(col / nx);
    const int i = a.d.i0 + (int)(col - (long long)jj * nx), j = a.d.j0 + jj;
    const long long pl = plane_index(a.d, i, j);
    const int k = bottom_k(a, pl);
    SedPoint c;
    gather(a, i, j, pl, k, c, nullptr);
    double out[OBM_SED_MAX_COUPLED] = {0, 0, 0, 0};
    coupled_fluxes(a.p, c, out);
    const double dz = a.d.zc[k] - a.d.zc[k - 1];  // Δzᶜᶜᶠ(i, j, k, grid)
    const long long idx = cell_index(a.d, i, j, k);
#pragma unroll
    for (int n = 0; n < OBM_SED_MAX_COUPLED; n++)
        if (n < a.nc && a.f.G_coupled[n]) a.f.G_coupled[n][idx] += out[n] / dz;  // K11
}

struct BottomArgs {
    GridDims d;
    const double* bottom_height;
    long long* out;
};
__global__ void __launch_bounds__(128) find_bottom_cell_kernel(const __grid_constant__ BottomArgs a) {
    const int nx = a.d.i1 - a.d.i0;
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= (long long)nx * (a.d.j1 - a.d.j0)) return;
    const int jj = (int)(col / nx);
    const long long pl = plane_index(a.d, a.d.i0 + (int)(col - (long long)jj * nx), a.d.j0 + jj);
    const double h = a.bottom_height[pl];
    int kb = 1;
    while ((a.d.zc[kb - 1] <= h) && (kb < a.d.Nz)) kb += 1;
    a.out[pl] = kb;
}

static int fill_args(SedArgs& a, const obm_grid* grid, const obm_sediment_params* p, const obm_sediment_fields* f,
                     bool need_step) {
    OBM_REQUIRE(p && f, OBM_ENULL, "sediment: params / fields is NULL");
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    OBM_REQUIRE(a.d.Hz >= 1, OBM_ESIZE, "sediment kernels need Hz >= 1 (they read the cell below the bottom cell)");
    OBM_REQUIRE(p->model == OBM_SED_INSTANT_REMINERALISATION || p->model == OBM_SED_SIMPLE_MULTI_G, OBM_EENUM,
                "sediment: unknown model %d", p->model);
    OBM_REQUIRE(p->nsinking_nitrogen >= 1 && p->nsinking_nitrogen <= OBM_SED_MAX_SINKING && p->nsinking_carbon >= 0
                    && p->nsinking_carbon <= OBM_SED_MAX_SINKING,
                OBM_ESIZE, "sediment: bad sinking tracer counts (%d, %d)", p->nsinking_nitrogen, p->nsinking_carbon);
    a.p = *p;
    a.f = *f;
    const bool smg = p->model == OBM_SED_SIMPLE_MULTI_G;
    a.np = smg ? (p->carbon ? 6 : 3) : 1;
    a.nc = smg ? (p->carbon ? 4 : 3) : 1;
    a.nt = smg ? 3 : 0;
    OBM_REQUIRE(!smg || (f->NO3 && f->NH4 && f->O2), OBM_ENULL, "sediment: NO3 / NH4 / O2 is NULL");
    OBM_REQUIRE(!(smg && p->carbon) || p->nsinking_carbon >= 1, OBM_ESIZE, "sediment: carbon variant needs sinking carbon tracers");
    for (int n = 0; n < p->nsinking_nitrogen + p->nsinking_carbon; n++)
        OBM_REQUIRE(f->sinking[n] && f->sinking_w[n], OBM_ENULL, "sediment: sinking tracer / velocity %d is NULL", n);
    for (int n = 0; n < a.np; n++) {
        OBM_REQUIRE(f->pools[n], OBM_ENULL, "sediment: pool %d is NULL", n);
        OBM_REQUIRE(!need_step || (f->Gn[n] && f->Gm[n]), OBM_ENULL, "sediment: Gn / Gm of pool %d is NULL", n);
    }
    return 0;
}

}  // namespace obm

using namespace obm;

extern "C" int obm_sediment_update_state(const obm_grid* grid, const obm_sediment_params* p, const obm_sediment_fields* f,
                                         double dt, double chi, void* stream) {
    SedArgs a;
    memset(&a, 0, sizeof(a));
    int rc = fill_args(a, grid, p, f, true);
    if (rc) return rc;
    a.dt = dt; a.chi = chi;
    const long long ncols = column_count(a.d);
    sediment_state_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("sediment_state_kernel");
}

extern "C" int obm_sediment_update_tendencies(const obm_grid* grid, const obm_sediment_params* p,
                                              const obm_sediment_fields* f, void* stream) {
    SedArgs a;
    memset(&a, 0, sizeof(a));
    int rc = fill_args(a, grid, p, f, false);
    if (rc) return rc;
    const long long ncols = column_count(a.d);
    sediment_tendency_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("sediment_tendency_kernel");
}

extern "C" int obm_find_bottom_cells(const obm_grid* grid, const double* bottom_height_xy, int64_t* bottom_indices_xy,
                                     void* stream) {
    OBM_REQUIRE(bottom_height_xy && bottom_indices_xy, OBM_ENULL, "obm_find_bottom_cells: a pointer is NULL");
    BottomArgs a;
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    a.bottom_height = bottom_height_xy;
    a.out = (long long*)bottom_indices_xy;
    const long long ncols = column_count(a.d);
    find_bottom_cell_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("find_bottom_cell_kernel");
}
