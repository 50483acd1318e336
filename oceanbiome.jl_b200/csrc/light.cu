// light.cu — column light-attenuation scans (two-band, N-band), euphotic depth, mixed-layer mean.
//
// Replaces the serial-in-z, one-thread-per-column kernels of
//   src/Light/2band.jl:1-33, src/Light/multi_band.jl:147-163 (one launch PER BAND in the
//   reference, :170-180), src/Light/compute_euphotic_depth.jl:3-29,
//   PISCES/mean_mixed_layer_properties.jl:25-49.
//
// Design (SURVEY §5 "long axis"): the only sequential axis is depth.  A block owns a tile of 32
// columns that are adjacent in memory (x fastest) and walks it top-down in z-tiles of 32 levels:
//   1. the 8 warps load the [32 levels × 32 columns] tile with fully coalesced 256-byte rows into
//      padded shared memory (the transpose buffer);
//   2. each warp then owns 4 columns; within a column lane ↔ level, so the 2 pow + 2 exp per
//      level (the whole FP64 cost) run 32-wide, and the serial recurrence becomes a warp
//      shuffle scan (prefix SUM of the chlorophyll integral for two-band, prefix PRODUCT of the
//      per-level transmittances for N-band) with the running value carried between z-tiles in a
//      register;
//   3. results go back through shared memory and are stored with coalesced rows.
// All bands of the N-band model are produced in one launch from one read of chlorophyll.
#include "obm_common.cuh"

namespace obm {

// exp of the scans: the lean exp of obm_common.cuh (≈ 26 instructions, ≤ 2 ulp, library call for |x| ≥ 700 / NaN) —
// these kernels are issue-bound on their 4 (two-band) / 2·bands (N-band) exps per cell: 3-band PAR 3.2 → 2.65 ms.
// -DOBM_LIGHT_EXP=0 restores the library exp.
#ifndef OBM_LIGHT_EXP
#define OBM_LIGHT_EXP 1
#endif
__device__ __forceinline__ double lexp(double x) {
#if OBM_LIGHT_EXP == 0
    return exp(x);
#else
    return exp_lean(x);
#endif
}

constexpr int TC = 32;        // columns per block tile
constexpr int TZ = 32;        // levels per z-tile (= warp width)
constexpr int NWARP = 8;      // warps per block
constexpr int CPW = TC / NWARP;  // columns per warp

__device__ __forceinline__ double warp_inclusive_sum(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
__device__ __forceinline__ double warp_inclusive_prod(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v *= n;
    }
    return v;
}

// column c of the block tile → parent offset of its k = 0 cell (or -1 when out of range)
__device__ __forceinline__ long long column_base(const GridDims& d, long long c, long long ncols, long long* plane) {
    if (c >= ncols) return -1;
    const int nx = d.i1 - d.i0;
    const int jj = (int)(c / nx);
    const int i = d.i0 + (int)(c - (long long)jj * nx);
    const int j = d.j0 + jj;
    *plane = plane_index(d, i, j);
    return cell_index(d, i, j, 0);
}

struct TwoBandArgs {
    GridDims d;
    obm_twoband_params m;
    const double* P;
    const double* sPAR;
    double sPAR_const;
    double* PAR;
};

__global__ void __launch_bounds__(TC* NWARP) par_twoband_kernel(const __grid_constant__ TwoBandArgs a) {
    __shared__ double tile[TZ][TC + 1];
    __shared__ long long col_base[TC];
    __shared__ double col_par0[TC];

    const GridDims& d = a.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ncols = (long long)(d.i1 - d.i0) * (d.j1 - d.j0);
    if (warp == 0) {
        long long plane = 0;
        long long b = column_base(d, (long long)blockIdx.x * TC + lane, ncols, &plane);
        col_base[lane] = b;
        col_par0[lane] = b < 0 ? 0.0 : (a.sPAR ? a.sPAR[plane] : a.sPAR_const);
    }
    __syncthreads();

    const double kr = a.m.water_red_attenuation, kb = a.m.water_blue_attenuation;
    const double xr = a.m.chlorophyll_red_attenuation, xb = a.m.chlorophyll_blue_attenuation;
    const double er = a.m.chlorophyll_red_exponent, eb = a.m.chlorophyll_blue_exponent;
    const double Rcp = a.m.phytoplankton_chlorophyll_ratio, r = a.m.pigment_ratio;
    const int Nz = d.Nz;

    // per-column carries (uniform across the warp): running ∫chl of both bands and the pigment
    // powers of the level just above the current z-tile
    double carry_r[CPW], carry_b[CPW], prev_pr[CPW], prev_pb[CPW];
#pragma unroll
    for (int q = 0; q < CPW; q++) carry_r[q] = carry_b[q] = prev_pr[q] = prev_pb[q] = 0.0;

    for (int ktop = Nz - 1; ktop >= 0; ktop -= TZ) {
        // 1. coalesced load: tile row l ↔ level ktop - l
#pragma unroll
        for (int q = 0; q < TZ / NWARP; q++) {
            const int l = warp * (TZ / NWARP) + q, k = ktop - l;
            const long long b = col_base[lane];
            tile[l][lane] = (k >= 0 && b >= 0) ? a.P[b + d.sz * k] : 0.0;
        }
        __syncthreads();
        // 2. lane ↔ level
        const int k = ktop - lane;
        const bool live = k >= 0;
        const double zck = live ? d.zc[k] : 0.0;
        // weights of 2band.jl:20-21 (top level) / :28-29 (the rest)
        const double w_above = (live && k < Nz - 1) ? (d.zc[k + 1] - d.zf[k + 1]) : 0.0;
        const double w_here = live ? (d.zf[k + 1] - zck) : 0.0;
#pragma unroll
        for (int q = 0; q < CPW; q++) {
            const int c = warp * CPW + q;
            // (P Rᶜₚ / r)^e for both bands from ONE logarithm: x^e = exp(e ln x)  (x = 0 → 0, x < 0 → NaN like pow;
            // |e ln x| ≲ 10 ⇒ ≤ 2e-15 relative, inside the 1e-12 tolerance; saves two ≈ 130-instruction pow calls)
            const double lp = log(tile[lane][c] * Rcp / r);
            const double pr = live ? (er == 0.0 ? 1.0 : lexp(er * lp)) : 0.0;  // x^0 ≡ 1
            const double pb = live ? (eb == 0.0 ? 1.0 : lexp(eb * lp)) : 0.0;
            double pr_up = __shfl_up_sync(0xffffffffu, pr, 1);
            double pb_up = __shfl_up_sync(0xffffffffu, pb, 1);
            if (lane == 0) { pr_up = prev_pr[q]; pb_up = prev_pb[q]; }
            // increment of the integral at this level; w_above = 0 at the top level so the
            // (undefined) level above contributes exactly +0
            const double dr = w_above * pr_up + w_here * pr;
            const double db = w_above * pb_up + w_here * pb;
            const double ir = carry_r[q] + warp_inclusive_sum(dr, lane);
            const double ib = carry_b[q] + warp_inclusive_sum(db, lane);
            const double par = col_par0[c] * (lexp(kr * zck - xr * ir) + lexp(kb * zck - xb * ib)) / 2;
            carry_r[q] = __shfl_sync(0xffffffffu, ir, 31);
            carry_b[q] = __shfl_sync(0xffffffffu, ib, 31);
            prev_pr[q] = __shfl_sync(0xffffffffu, pr, 31);
            prev_pb[q] = __shfl_sync(0xffffffffu, pb, 31);
            tile[lane][c] = par;  // in place: only this warp touches column c between the barriers
        }
        __syncthreads();
        // 3. coalesced store
#pragma unroll
        for (int q = 0; q < TZ / NWARP; q++) {
            const int l = warp * (TZ / NWARP) + q, kk = ktop - l;
            const long long b = col_base[lane];
            if (kk >= 0 && b >= 0) a.PAR[b + d.sz * kk] = tile[l][lane];
        }
        __syncthreads();
    }
}

struct MultiBandArgs {
    GridDims d;
    obm_multiband_params m;
    const double* chl_a;
    const double* chl_b;
    double chl_scale;
    const double* sPAR;
    double sPAR_const;
    double* bands[OBM_MAX_BANDS];
    double* total;
};

// NB = number of bands (compile-time so the per-band carries live in registers)
template <int NB>
__global__ void __launch_bounds__(TC* NWARP) par_multiband_kernel(const __grid_constant__ MultiBandArgs a) {
    __shared__ double tile[TZ][TC + 1];
    __shared__ double out[NB][TZ][TC + 1];
    __shared__ long long col_base[TC];
    __shared__ double col_par0[TC];

    const GridDims& d = a.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ncols = (long long)(d.i1 - d.i0) * (d.j1 - d.j0);
    if (warp == 0) {
        long long plane = 0;
        long long b = column_base(d, (long long)blockIdx.x * TC + lane, ncols, &plane);
        col_base[lane] = b;
        col_par0[lane] = b < 0 ? 0.0 : (a.sPAR ? a.sPAR[plane] : a.sPAR_const);
    }
    __syncthreads();
    const int Nz = d.Nz;

    double carry[CPW][NB];  // field[k+1] of the level just above the z-tile, per column and band
#pragma unroll
    for (int q = 0; q < CPW; q++)
#pragma unroll
        for (int n = 0; n < NB; n++) carry[q][n] = 1.0;

    for (int ktop = Nz - 1; ktop >= 0; ktop -= TZ) {
#pragma unroll
        for (int q = 0; q < TZ / NWARP; q++) {
            const int l = warp * (TZ / NWARP) + q, k = ktop - l;
            const long long b = col_base[lane];
            double v = 0.0;
            if (k >= 0 && b >= 0) {
                const long long idx = b + d.sz * k;
                v = a.chl_b ? a.chl_scale * (a.chl_a[idx] + a.chl_b[idx]) : a.chl_scale * a.chl_a[idx];
            }
            tile[l][lane] = v;
        }
        __syncthreads();
        const int k = ktop - lane;
        const bool live = k >= 0;
        // multi_band.jl:156 (k = Nz: factor zᶜ[Nz], seeded with surface_PAR·division) / :160-161 (Δz)
        const double dz = !live ? 0.0 : (k == Nz - 1 ? d.zc[k] : d.zc[k] - d.zc[k + 1]);
#pragma unroll
        for (int q = 0; q < CPW; q++) {
            const int c = warp * CPW + q;
            const double lchl = log(tile[lane][c]);  // Chl^e = exp(e ln Chl): one log shared by all bands
#pragma unroll
            for (int n = 0; n < NB; n++) {
                const double kw = a.m.water_attenuation_coefficient[n], e = a.m.chlorophyll_exponent[n];
                const double chi = a.m.chlorophyll_attenuation_coefficient[n];
                const double chle = e == 0.0 ? 1.0 : lexp(e * lchl);  // x^0 ≡ 1 (also for x = 0)
                double t = live ? lexp(dz * (kw + chi * chle)) : 1.0;
                if (ktop == Nz - 1 && lane == 0) t = col_par0[c] * a.m.surface_PAR_division[n] * t;
                const double f = carry[q][n] * warp_inclusive_prod(t, lane);
                carry[q][n] = __shfl_sync(0xffffffffu, f, 31);
                out[n][lane][c] = f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < TZ / NWARP; q++) {
            const int l = warp * (TZ / NWARP) + q, kk = ktop - l;
            const long long b = col_base[lane];
            if (kk >= 0 && b >= 0) {
                const long long idx = b + d.sz * kk;
                double s = 0.0;
#pragma unroll
                for (int n = 0; n < NB; n++) {
                    const double f = out[n][l][lane];
                    if (a.bands[n]) a.bands[n][idx] = f;
                    s = (n == 0) ? f : s + f;  // sum(fields): ((PAR₁ + PAR₂) + PAR₃) multi_band.jl:120
                }
                if (a.total) a.total[idx] = s;
            }
        }
        __syncthreads();
    }
}

// ---- euphotic depth: compute_euphotic_depth.jl:3-29 — one thread per column, coalesced in x ----
struct ZeuArgs {
    GridDims d;
    const double* PAR;
    double cutoff;
    double* zeu;
};
__global__ void __launch_bounds__(128) euphotic_depth_kernel(const __grid_constant__ ZeuArgs a) {
    const GridDims& d = a.d;
    const long long ncols = (long long)(d.i1 - d.i0) * (d.j1 - d.j0);
    long long plane = 0;
    const long long b = column_base(d, (long long)blockIdx.x * blockDim.x + threadIdx.x, ncols, &plane);
    if (b < 0) return;
    const int Nz = d.Nz;
    // surface value uses the (unfilled) top halo cell exactly as the reference does (:6)
    const double surface = (a.PAR[b + d.sz * (Nz - 1)] + a.PAR[b + d.sz * Nz]) / 2;
    const double thr = surface * a.cutoff;
    double zeu = -INFINITY;
    double above = a.PAR[b + d.sz * (Nz - 1)];
    for (int k = Nz - 2; k >= 0; k--) {
        const double here = a.PAR[b + d.sz * k];
        if ((here <= thr) && isinf(zeu)) {
            const double zk = d.zc[k], zk1 = d.zc[k + 1];
            zeu = zk + (log(thr) - log(here)) * (zk - zk1) / (log(here) - log(above));
        }
        above = here;
    }
    a.zeu[plane] = isfinite(zeu) ? zeu : d.zc[-1];  // znode(i, j, 0, grid, …) :28
}

// ---- mixed-layer mean: PISCES/mean_mixed_layer_properties.jl:25-49 ---------------------------
struct MlmArgs {
    GridDims d;
    const double* zmxl;
    const double* C;
    double C_const;  // used when C == nullptr (ConstantField)
    double* out;
};
__global__ void __launch_bounds__(128) mixed_layer_mean_kernel(const __grid_constant__ MlmArgs a) {
    const GridDims& d = a.d;
    const long long ncols = (long long)(d.i1 - d.i0) * (d.j1 - d.j0);
    long long plane = 0;
    const long long b = column_base(d, (long long)blockIdx.x * blockDim.x + threadIdx.x, ncols, &plane);
    if (b < 0) return;
    const double zmxl = a.zmxl[plane];
    double acc = 0.0, depth = 0.0;
    for (int k = d.Nz - 1; k >= 0; k--) {
        const double zk = d.zf[k], zk1 = d.zf[k + 1];
        const double dzk = zk1 - zk;
        const double dzk1 = zk1 > zmxl ? zk1 - zmxl : 0.0;
        const double dz = zk >= zmxl ? dzk : dzk1;
        const double c = a.C ? a.C[b + d.sz * k] : a.C_const;
        acc += c * dz;
        depth += dz;
    }
    a.out[plane] = acc / depth;
}

}  // namespace obm

using namespace obm;

extern "C" int obm_par_twoband(const obm_grid* grid, const obm_twoband_params* p, const double* P,
                               const double* surface_PAR_xy, double surface_PAR_const, double* PAR, void* stream) {
    OBM_REQUIRE(p && P && PAR, OBM_ENULL, "obm_par_twoband: params / P / PAR is NULL");
    TwoBandArgs a;
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    a.m = *p;
    a.P = P;
    a.sPAR = surface_PAR_xy;
    a.sPAR_const = surface_PAR_const;
    a.PAR = PAR;
    const long long ncols = column_count(a.d);
    par_twoband_kernel<<<(unsigned)((ncols + TC - 1) / TC), TC * NWARP, 0, (cudaStream_t)stream>>>(a);
    return launch_status("par_twoband_kernel");
}

extern "C" int obm_par_multiband(const obm_grid* grid, const obm_multiband_params* p, const double* chl_a,
                                 const double* chl_b, double chl_scale, const double* surface_PAR_xy,
                                 double surface_PAR_const, double* const* PAR_bands, double* PAR_total, void* stream) {
    OBM_REQUIRE(p && chl_a && PAR_bands, OBM_ENULL, "obm_par_multiband: params / chl_a / PAR_bands is NULL");
    OBM_REQUIRE(p->nbands >= 1 && p->nbands <= 4, p->nbands >= 1 && p->nbands <= OBM_MAX_BANDS ? OBM_ENOTIMPL : OBM_ESIZE,
                "obm_par_multiband: nbands = %d (this build supports 1..4)", p->nbands);
    MultiBandArgs a;
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    a.m = *p;
    a.chl_a = chl_a;
    a.chl_b = chl_b;
    a.chl_scale = chl_scale;
    a.sPAR = surface_PAR_xy;
    a.sPAR_const = surface_PAR_const;
    for (int n = 0; n < OBM_MAX_BANDS; n++) a.bands[n] = n < p->nbands ? PAR_bands[n] : nullptr;
    a.total = PAR_total;
    const long long ncols = column_count(a.d);
    const unsigned blocks = (unsigned)((ncols + TC - 1) / TC);
    cudaStream_t s = (cudaStream_t)stream;
    switch (p->nbands) {
        case 1: par_multiband_kernel<1><<<blocks, TC * NWARP, 0, s>>>(a); break;
        case 2: par_multiband_kernel<2><<<blocks, TC * NWARP, 0, s>>>(a); break;
        case 3: par_multiband_kernel<3><<<blocks, TC * NWARP, 0, s>>>(a); break;
        default: par_multiband_kernel<4><<<blocks, TC * NWARP, 0, s>>>(a); break;
    }
    return launch_status("par_multiband_kernel");
}

extern "C" int obm_euphotic_depth(const obm_grid* grid, const double* PAR, double cutoff, double* zeu_xy, void* stream) {
    OBM_REQUIRE(PAR && zeu_xy, OBM_ENULL, "obm_euphotic_depth: PAR / zeu is NULL");
    ZeuArgs a;
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    OBM_REQUIRE(a.d.Hz >= 1, OBM_ESIZE, "obm_euphotic_depth needs Hz >= 1 (reads PAR[i,j,Nz+1] and znode(k=0))");
    a.PAR = PAR;
    a.cutoff = cutoff;
    a.zeu = zeu_xy;
    const long long ncols = column_count(a.d);
    euphotic_depth_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("euphotic_depth_kernel");
}

extern "C" int obm_mixed_layer_mean(const obm_grid* grid, const double* mixed_layer_depth_xy, const double* C,
                                    double C_const, double* mean_xy, void* stream) {
    OBM_REQUIRE(mixed_layer_depth_xy && mean_xy, OBM_ENULL, "obm_mixed_layer_mean: zmxl / out is NULL");
    MlmArgs a;
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    a.zmxl = mixed_layer_depth_xy;
    a.C = C;
    a.C_const = C_const;
    a.out = mean_xy;
    const long long ncols = column_count(a.d);
    mixed_layer_mean_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("mixed_layer_mean_kernel");
}
