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
// -DOBM_LIGHT_EXP=0 restores the library exp; 2 / 3 / 4: see lexp (4, the table form, is the fastest: profiles/r03_kernel_variants.txt).
#ifndef OBM_LIGHT_EXP
#define OBM_LIGHT_EXP 4
#endif
// OBM_LIGHT_LOG: 1 = the lean logarithm of obm_common.cuh for the four levels of a lane, branch-free in one block with ONE
// combined range test (a non-positive, subnormal, infinite or NaN argument sends all four through the library log);
// 0 = the library log.
#ifndef OBM_LIGHT_LOG
#define OBM_LIGHT_LOG 1
#endif
__device__ __forceinline__ void log4(const double (&x)[4], double (&out)[4]) {
#if OBM_LIGHT_LOG
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        out[i] = log_unguarded(x[i]);
        ok &= log_in_range(x[i]);
    }
    if (!ok) {
#pragma unroll
        for (int i = 0; i < 4; i++) out[i] = log(x[i]);
    }
#else
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = log(x[i]);
#endif
}
// OBM_PAR_SCAN4: four levels per lane in the scans (see par_multiband_kernel): 3-band PAR 0.392 → 0.350 ms per 16.8 M cells
// (profiles/r04_kernel_variants.txt); 0 restores one level per lane and a 5-step scan per level.
#ifndef OBM_PAR_SCAN4
#define OBM_PAR_SCAN4 1
#endif
__device__ __forceinline__ double lexp(double x) {
#if OBM_LIGHT_EXP == 0
    return exp(x);
#elif OBM_LIGHT_EXP == 1
    return exp_lean(x);
#elif OBM_LIGHT_EXP == 2
    return exp_horner(x);   // plain Horner, integer range test, library exp inline in the cold branch (no call)
#elif OBM_LIGHT_EXP == 4
    return exp_table_clamped(x);  // 64-entry table + degree-5 polynomial (obm_common.cuh), same clamp semantics as 3
#else
    // branch-free: the argument is clamped to where the exponent-field arithmetic of exp_unguarded is valid (normal
    // results).  −708 → 3.3e-308 instead of the subnormals / exact 0 below it (Chl = 0: ln 0 = −Inf and χ·3e-308 vanishes
    // in kʷ + χ Chl^e; an attenuation factor that small multiplies a PAR already far below any threshold), 709 → 8e307
    // instead of +Inf (Chl ≳ 1e300); NaN stays NaN.
    return exp_unguarded(x < -708.0 ? -708.0 : (x > 709.0 ? 709.0 : x));
#endif
}

// OBM_LIGHT_EXP_BATCH (with the table exp): the exps of a lane's four levels run unclamped, with ONE combined range test
// (an integer compare per argument instead of the clamp's two FP64 compares and two 64-bit selects); a lane with an
// argument outside |x| < 700 — Chl = 0 (ln 0 = −Inf), a level below the bottom, NaN — redoes its four levels with the
// clamped form: same results, the cold branch is out of the way of the other 99.9 %.
// Timed (profiles/r04_kernel_variants.txt, visit r4l): +2.5 % on the 3-band scan, +7 % on the two-band one — the fallback's live values
// cost 16 – 48 B more stack than the clamps' selects cost issue slots — so it is OFF; kept as a build option.
#ifndef OBM_LIGHT_EXP_BATCH
#define OBM_LIGHT_EXP_BATCH 0
#endif
__device__ __forceinline__ double lexp_try(double x, bool& ok) {
#if OBM_LIGHT_EXP_BATCH
    ok &= exp_in_range(x);
    return exp_table(x);
#else
    return lexp(x);
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

#ifndef OBM_PAR_TWOBAND_BLOCKS
#define OBM_PAR_TWOBAND_BLOCKS 4  // 64 registers; 3 (≤ 80): +3 %, 5 (48): +5 % (profiles/r04_kernel_variants.txt)
#endif
__global__ void __launch_bounds__(TC* NWARP, OBM_PAR_TWOBAND_BLOCKS) par_twoband_kernel(const __grid_constant__ TwoBandArgs a) {
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
#if OBM_PAR_SCAN4
        // 2. lane ↔ (column q = lane & 3 of the warp's four, levels 4g … 4g + 3 of the z-tile, g = lane >> 2): three serial
        // additions inside the lane, a 3-step scan over the column's eight lanes (stride 4), one addition per level —
        // see par_multiband_kernel.  The level above a lane's first level is the last level of lane − 4 (or the carry).
        {
            const int q4 = lane & 3, g = lane >> 2;
            const int c = warp * CPW + q4;
            double pr[4], pb[4], zck[4], w_above[4], w_here[4], lx[4], lp4[4];
            bool live4[4];
#pragma unroll
            for (int i = 0; i < 4; i++) lx[i] = tile[4 * g + i][c] * Rcp / r;
            log4(lx, lp4);  // x^e = exp(e ln x), one logarithm for both bands
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = ktop - (4 * g + i);
                const bool live = k >= 0;
                zck[i] = live ? d.zc[k] : 0.0;
                // weights of 2band.jl:20-21 (top level) / :28-29 (the rest)
                w_above[i] = (live && k < Nz - 1) ? (d.zc[k + 1] - d.zf[k + 1]) : 0.0;
                w_here[i] = live ? (d.zf[k + 1] - zck[i]) : 0.0;
                live4[i] = live;
            }
            {
                bool ok = true;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    pr[i] = live4[i] ? (er == 0.0 ? 1.0 : lexp_try(er * lp4[i], ok)) : 0.0;  // x^0 ≡ 1
                    pb[i] = live4[i] ? (eb == 0.0 ? 1.0 : lexp_try(eb * lp4[i], ok)) : 0.0;
                }
                if (!ok) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        pr[i] = live4[i] ? (er == 0.0 ? 1.0 : lexp(er * lp4[i])) : 0.0;
                        pb[i] = live4[i] ? (eb == 0.0 ? 1.0 : lexp(eb * lp4[i])) : 0.0;
                    }
                }
            }
            double pr_up = __shfl_up_sync(0xffffffffu, pr[3], 4);
            double pb_up = __shfl_up_sync(0xffffffffu, pb[3], 4);
            if (g == 0) { pr_up = prev_pr[0]; pb_up = prev_pb[0]; }
            // running integral inside the lane; w_above = 0 at the top level so the (undefined) level above adds exactly +0
            double sr[4], sb[4];
            sr[0] = w_above[0] * pr_up + w_here[0] * pr[0];
            sb[0] = w_above[0] * pb_up + w_here[0] * pb[0];
#pragma unroll
            for (int i = 1; i < 4; i++) {
                sr[i] = sr[i - 1] + (w_above[i] * pr[i - 1] + w_here[i] * pr[i]);
                sb[i] = sb[i - 1] + (w_above[i] * pb[i - 1] + w_here[i] * pb[i]);
            }
            double scr = sr[3], scb = sb[3];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const double nr = __shfl_up_sync(0xffffffffu, scr, 4 * o);
                const double nb = __shfl_up_sync(0xffffffffu, scb, 4 * o);
                if (g >= o) { scr += nr; scb += nb; }
            }
            double exr = __shfl_up_sync(0xffffffffu, scr, 4);
            double exb = __shfl_up_sync(0xffffffffu, scb, 4);
            if (g == 0) exr = exb = 0.0;
            const double base_r = carry_r[0] + exr, base_b = carry_b[0] + exb;
            const double par0 = col_par0[c];
            const double ir3 = base_r + sr[3], ib3 = base_b + sb[3];
            {
                bool ok = true;
                double par4[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double ir = base_r + sr[i], ib = base_b + sb[i];
                    par4[i] = par0 * (lexp_try(kr * zck[i] - xr * ir, ok) + lexp_try(kb * zck[i] - xb * ib, ok)) / 2;
                }
                if (!ok) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const double ir = base_r + sr[i], ib = base_b + sb[i];
                        par4[i] = par0 * (lexp(kr * zck[i] - xr * ir) + lexp(kb * zck[i] - xb * ib)) / 2;
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; i++) tile[4 * g + i][c] = par4[i];  // in place
            }
            carry_r[0] = __shfl_sync(0xffffffffu, ir3, 28 + q4);
            carry_b[0] = __shfl_sync(0xffffffffu, ib3, 28 + q4);
            prev_pr[0] = __shfl_sync(0xffffffffu, pr[3], 28 + q4);
            prev_pb[0] = __shfl_sync(0xffffffffu, pb[3], 28 + q4);
        }
#else
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
#endif
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
    // DIAG variant: the two column diagnostics PISCES derives from the total PAR (PISCES/update_state.jl:7,11)
    const double* zmxl;  // 2-D mixed-layer depth
    double cutoff;       // euphotic depth = where PAR falls to cutoff × surface PAR
    double* zeu;         // 2-D out
    double* mlmean;      // 2-D out: mixed-layer mean of the total PAR
};

#ifndef OBM_PAR_DIAG_BLOCKS
#define OBM_PAR_DIAG_BLOCKS 5
#endif
// compute_euphotic_depth.jl:20: log-space interpolation between level k (`here`, at or below the threshold) and the
// level above it.  Out of line: it runs a few times per column and carries four logs.
__device__ __noinline__ double euphotic_depth_between(double zk, double zk1, double here, double above, double thr) {
    return zk + (log(thr) - log(here)) * (zk - zk1) / (log(here) - log(above));
}

// NB = number of bands (compile-time so the per-band carries live in registers).
// DIAG: the launch also produces PISCES' euphotic depth (compute_euphotic_depth.jl:3-29) and mixed-layer mean PAR
// (mean_mixed_layer_properties.jl:25-49) of each column while the total PAR of a level is on chip — the two extra column
// launches and their re-read of the PAR field go away.  Both ride in the store phase (thread ↔ column, warp ↔ 4 levels):
// a thread adds its levels to its own partial sums (reduced across the 8 warps once, at the end: pairwise instead of the
// reference's top-down order, ≤ 1e-15 relative apart), flags the first level of the tile at or below the threshold
// with a shared-memory atomicMin, and warp 0 resolves the flagged level (log-space interpolation with the level above)
// between two barriers that are there anyway.
template <int NB, bool DIAG>
__global__ void __launch_bounds__(TC* NWARP, DIAG ? OBM_PAR_DIAG_BLOCKS : 5) par_multiband_kernel(const __grid_constant__ MultiBandArgs a) {
    __shared__ double tile[TZ][TC + 1];
    __shared__ double out[NB][TZ][TC + 1];
    __shared__ long long col_base[TC];
    __shared__ double col_par0[TC];
    constexpr int DT = DIAG ? TC : 1;
    __shared__ long long col_plane[DT];
    __shared__ double col_zmxl[DT], col_halo[DT], col_thr[DT], col_prev[DT], col_zeu[DT];
    __shared__ int col_first[DT], col_found[DT];

    const GridDims& d = a.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ncols = (long long)(d.i1 - d.i0) * (d.j1 - d.j0);
    if (warp == 0) {
        long long plane = 0;
        long long b = column_base(d, (long long)blockIdx.x * TC + lane, ncols, &plane);
        col_base[lane] = b;
        col_par0[lane] = b < 0 ? 0.0 : (a.sPAR ? a.sPAR[plane] : a.sPAR_const);
        if (DIAG) {
            col_plane[lane] = plane;
            col_zmxl[lane] = b < 0 ? 0.0 : a.zmxl[plane];
            // PAR[i, j, Nz + 1]: the (unfilled) top halo cell, read as found — compute_euphotic_depth.jl:6
            col_halo[lane] = b < 0 ? 0.0 : a.total[b + d.sz * d.Nz];
            col_thr[lane] = col_prev[lane] = 0.0;
            col_zeu[lane] = -INFINITY;
            col_first[lane] = TZ;
            col_found[lane] = 0;
        }
    }
    __syncthreads();
    const int Nz = d.Nz;

    double carry[CPW][NB];  // field[k+1] of the level just above the z-tile, per column and band
#pragma unroll
    for (int q = 0; q < CPW; q++)
#pragma unroll
        for (int n = 0; n < NB; n++) carry[q][n] = 1.0;
    double ml_acc = 0.0, ml_depth = 0.0;  // DIAG: this thread's share of Σ PAR·Δz and Σ Δz above zₘₓₗ (column = lane)

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
#if OBM_PAR_SCAN4
        // lane ↔ (column q = lane & 3 of the warp's four, levels 4g … 4g + 3 of the z-tile, g = lane >> 2): the prefix product
        // is three serial multiplications inside the lane, a 3-step scan over the eight lanes of the column (stride 4) and
        // one multiplication per level — 7 issue slots per cell and band instead of the 27 of a 5-step scan per level.
        // Shared-memory rows 4g + i at column 4·warp + q: (4g + q) mod 16 is distinct over a half-warp — conflict-free.
        {
            const int q4 = lane & 3, g = lane >> 2;
            const int c = warp * CPW + q4;
            double lchl[4], dz[4], chl4[4];
            bool live[4];
#pragma unroll
            for (int i = 0; i < 4; i++) chl4[i] = tile[4 * g + i][c];
            log4(chl4, lchl);  // Chl^e = exp(e ln Chl): one log shared by all bands
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = ktop - (4 * g + i);
                live[i] = k >= 0;
                // multi_band.jl:156 (k = Nz: factor zᶜ[Nz], seeded with surface_PAR·division) / :160-161 (Δz)
                dz[i] = !live[i] ? 0.0 : (k == Nz - 1 ? d.zc[k] : d.zc[k] - d.zc[k + 1]);
            }
            double top = 0.0;
#pragma unroll
            for (int n = 0; n < NB; n++) {
                const double kw = a.m.water_attenuation_coefficient[n], e = a.m.chlorophyll_exponent[n];
                const double chi = a.m.chlorophyll_attenuation_coefficient[n];
                double t[4];
                bool ok = true;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double chle = e == 0.0 ? 1.0 : lexp_try(e * lchl[i], ok);  // x^0 ≡ 1 (also for x = 0)
                    t[i] = live[i] ? lexp_try(dz[i] * (kw + chi * chle), ok) : 1.0;
                }
                if (!ok) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const double chle = e == 0.0 ? 1.0 : lexp(e * lchl[i]);
                        t[i] = live[i] ? lexp(dz[i] * (kw + chi * chle)) : 1.0;
                    }
                }
                if (ktop == Nz - 1 && g == 0) t[0] = col_par0[c] * a.m.surface_PAR_division[n] * t[0];
                const double p1 = t[0] * t[1], p2 = p1 * t[2], p3 = p2 * t[3];
                double sc = p3;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    const double nb = __shfl_up_sync(0xffffffffu, sc, 4 * o);
                    if (g >= o) sc *= nb;
                }
                double ex = __shfl_up_sync(0xffffffffu, sc, 4);
                if (g == 0) ex = 1.0;
                const double base = carry[0][n] * ex;  // carry[0][n]: this lane's column (field[k+1] just above the z-tile)
                const double f0 = base * t[0], f1 = base * p1, f2 = base * p2, f3 = base * p3;
                carry[0][n] = __shfl_sync(0xffffffffu, f3, 28 + q4);
                out[n][4 * g + 0][c] = f0;
                out[n][4 * g + 1][c] = f1;
                out[n][4 * g + 2][c] = f2;
                out[n][4 * g + 3][c] = f3;
                if (DIAG) top = (n == 0) ? f0 : top + f0;
            }
            // threshold of the column from its top level: PAR[Nz] (+ the halo cell above it) — compute_euphotic_depth.jl:6
            if (DIAG && ktop == Nz - 1 && g == 0) col_thr[c] = (top + col_halo[c]) / 2 * a.cutoff;
        }
#else
        const int k = ktop - lane;
        const bool live = k >= 0;
        // multi_band.jl:156 (k = Nz: factor zᶜ[Nz], seeded with surface_PAR·division) / :160-161 (Δz)
        const double dz = !live ? 0.0 : (k == Nz - 1 ? d.zc[k] : d.zc[k] - d.zc[k + 1]);
#pragma unroll
        for (int q = 0; q < CPW; q++) {
            const int c = warp * CPW + q;
            const double lchl = log(tile[lane][c]);  // Chl^e = exp(e ln Chl): one log shared by all bands
            double top = 0.0;
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
                if (DIAG) top = (n == 0) ? f : top + f;
            }
            // threshold of the column from its top level: PAR[Nz] (+ the halo cell above it) — compute_euphotic_depth.jl:6
            if (DIAG && ktop == Nz - 1 && lane == 0) col_thr[c] = (top + col_halo[c]) / 2 * a.cutoff;
        }
#endif
        __syncthreads();
#pragma unroll
        for (int q = 0; q < TZ / NWARP; q++) {
            const int l = warp * (TZ / NWARP) + q, kk = ktop - l;
            const long long b = col_base[lane];
            double s = 0.0;
            if (kk >= 0 && b >= 0) {
                const long long idx = b + d.sz * kk;
#pragma unroll
                for (int n = 0; n < NB; n++) {
                    const double f = out[n][l][lane];
                    if (a.bands[n]) a.bands[n][idx] = f;
                    s = (n == 0) ? f : s + f;  // sum(fields): ((PAR₁ + PAR₂) + PAR₃) multi_band.jl:120
                }
                if (a.total) a.total[idx] = s;
                if (DIAG) {
                    const double zk = d.zf[kk], zk1 = d.zf[kk + 1], zm = col_zmxl[lane];  // mean_mixed_layer_properties.jl:33-45
                    const double dzk1 = zk1 > zm ? zk1 - zm : 0.0;
                    const double dzm = zk >= zm ? zk1 - zk : dzk1;
                    ml_acc += s * dzm;
                    ml_depth += dzm;
                    if (kk <= Nz - 2 && s <= col_thr[lane]) atomicMin(&col_first[lane], l);
                }
            }
            // the total PAR of (l, lane) replaces band 0 of the same element, which only this thread reads in this phase
            if (DIAG) out[0][l][lane] = s;
        }
        __syncthreads();
        if (DIAG && warp == 0) {  // compute_euphotic_depth.jl:14-24 for the levels of this tile, column = lane
            const double thr = col_thr[lane];
            if (!col_found[lane]) {
                for (int l = col_first[lane]; l < TZ; l++) {
                    const int kk = ktop - l;
                    if (kk < 0) break;
                    const double here = out[0][l][lane];
                    if (!(kk <= Nz - 2 && here <= thr)) continue;
                    const double above = l > 0 ? out[0][l - 1][lane] : col_prev[lane];
                    const double z = euphotic_depth_between(d.zc[kk], d.zc[kk + 1], here, above, thr);
                    if (!isinf(z)) {  // `isinf(zₑᵤ)` keeps the search going after an infinite candidate
                        col_zeu[lane] = z;
                        col_found[lane] = 1;
                        break;
                    }
                }
            }
            col_first[lane] = TZ;
            col_prev[lane] = out[0][TZ - 1][lane];
        }
    }
    if (DIAG) {
        // Σ over the 8 warps' partial sums per column, fixed order; `out` is free after the last barrier of the loop
        __syncthreads();  // warp 0 has finished reading the last tile's totals
        double (*red)[2][TC] = reinterpret_cast<double (*)[2][TC]>(&out[0][0][0]);
        red[warp][0][lane] = ml_acc;
        red[warp][1][lane] = ml_depth;
        __syncthreads();
        if (warp == 0 && col_base[lane] >= 0) {
            double acc = 0.0, dep = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; w++) { acc += red[w][0][lane]; dep += red[w][1][lane]; }
            a.mlmean[col_plane[lane]] = acc / dep;
            const double z = col_zeu[lane];
            a.zeu[col_plane[lane]] = isfinite(z) ? z : d.zc[-1];  // znode(i, j, 0, grid, …) :28
        }
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

static int launch_multiband(const char* who, const obm_grid* grid, const obm_multiband_params* p, const double* chl_a,
                            const double* chl_b, double chl_scale, const double* surface_PAR_xy, double surface_PAR_const,
                            double* const* PAR_bands, double* PAR_total, const double* zmxl, double cutoff, double* zeu,
                            double* mlmean, bool diag, void* stream) {
    OBM_REQUIRE(p && chl_a && PAR_bands, OBM_ENULL, "%s: params / chl_a / PAR_bands is NULL", who);
    OBM_REQUIRE(p->nbands >= 1 && p->nbands <= 4, p->nbands >= 1 && p->nbands <= OBM_MAX_BANDS ? OBM_ENOTIMPL : OBM_ESIZE,
                "%s: nbands = %d (this build supports 1..4)", who, p->nbands);
    MultiBandArgs a;
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    if (diag) {
        OBM_REQUIRE(PAR_total && zmxl && zeu && mlmean, OBM_ENULL, "%s: PAR_total / zmxl / zeu / mean is NULL", who);
        OBM_REQUIRE(a.d.Hz >= 1, OBM_ESIZE, "%s needs Hz >= 1 (reads PAR[i,j,Nz+1] and znode(k=0))", who);
    }
    a.m = *p;
    a.chl_a = chl_a;
    a.chl_b = chl_b;
    a.chl_scale = chl_scale;
    a.sPAR = surface_PAR_xy;
    a.sPAR_const = surface_PAR_const;
    for (int n = 0; n < OBM_MAX_BANDS; n++) a.bands[n] = n < p->nbands ? PAR_bands[n] : nullptr;
    a.total = PAR_total;
    a.zmxl = zmxl; a.cutoff = cutoff; a.zeu = zeu; a.mlmean = mlmean;
    const long long ncols = column_count(a.d);
    const unsigned blocks = (unsigned)((ncols + TC - 1) / TC);
    cudaStream_t s = (cudaStream_t)stream;
#define OBM_MB(NB)                                                                   \
    if (diag) par_multiband_kernel<NB, true><<<blocks, TC * NWARP, 0, s>>>(a);       \
    else par_multiband_kernel<NB, false><<<blocks, TC * NWARP, 0, s>>>(a)
    switch (p->nbands) {
        case 1: OBM_MB(1); break;
        case 2: OBM_MB(2); break;
        case 3: OBM_MB(3); break;
        default: OBM_MB(4); break;
    }
#undef OBM_MB
    return launch_status("par_multiband_kernel");
}

extern "C" int obm_par_multiband(const obm_grid* grid, const obm_multiband_params* p, const double* chl_a,
                                 const double* chl_b, double chl_scale, const double* surface_PAR_xy,
                                 double surface_PAR_const, double* const* PAR_bands, double* PAR_total, void* stream) {
    return launch_multiband("obm_par_multiband", grid, p, chl_a, chl_b, chl_scale, surface_PAR_xy, surface_PAR_const, PAR_bands,
                            PAR_total, nullptr, 0.0, nullptr, nullptr, false, stream);
}

extern "C" int obm_par_multiband_column_state(const obm_grid* grid, const obm_multiband_params* p, const double* chl_a,
                                              const double* chl_b, double chl_scale, const double* surface_PAR_xy,
                                              double surface_PAR_const, double* const* PAR_bands, double* PAR_total,
                                              const double* mixed_layer_depth_xy, double cutoff, double* zeu_xy,
                                              double* mean_mixed_layer_PAR_xy, void* stream) {
    return launch_multiband("obm_par_multiband_column_state", grid, p, chl_a, chl_b, chl_scale, surface_PAR_xy, surface_PAR_const,
                            PAR_bands, PAR_total, mixed_layer_depth_xy, cutoff, zeu_xy, mean_mixed_layer_PAR_xy, true, stream);
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
