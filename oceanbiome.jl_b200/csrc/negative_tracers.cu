// negative_tracers.cu — ScaleNegativeTracers (all conserved groups of a model in ONE launch),
// ZeroNegativeTracers, and the fused tracer-inventory reduction.
//
// Replaces src/Utils/negative_tracers.jl:137-276: the reference launches one :xyz kernel per
// conserved group (PISCES: 5 launches re-reading 36 fields; OceanBioME.jl:169).  Here a cell's
// distinct tracers are read once into shared memory (one 8-byte slot per tracer per thread,
// conflict-free), the groups are applied back to back on those staged values in the reference's
// order, and every tracer is written once.  HBM-bound: 16 B per distinct tracer per cell.
//
// The sums t, p use explicit round-to-nearest mul/add (no FMA contraction) — bit-identical to the
// reference's `t += value * scale` — so the decisions (t < 0, value > 0, NaN fill, zeroing) are exact.
// The rescaling itself uses ONE division per group (value·(t/p) instead of value·t/p, ≤ 1 ulp apart) and
// cells whose group has nothing to rescale are neither recomputed nor written back (the reference rewrites
// them with value·t/t, i.e. value up to 1 ulp of rounding noise): half the HBM traffic in the common case.
#include <string.h>

#include <atomic>

#include "carbon_chemistry.cuh"
#include "obm_common.cuh"

namespace obm {

constexpr int SN_BLOCK = 128;

struct ScaleArgs {
    GridDims d;
    int ntracers, ngroups;
    double fill;
    double* tracers[OBM_MAX_SCALE_TRACERS];
    obm_scale_group groups[OBM_MAX_SCALE_GROUPS];
};

// One cell: stage its distinct tracers in `mine` (stride SN_BLOCK), apply the groups in order, write back what changed.
// On return `mine` holds the cell's tracers as the rest of the stage will see them.
// STAGED: the values are already in `mine` (the caller's cp.async copies have landed).
template <bool STAGED = false>
__device__ __forceinline__ void scale_cell(const ScaleArgs& a, double* mine, long long idx) {
    // some value is negative or non-finite: only then can any group have p ≠ t.
    // negative, −0.0, ±Inf or NaN ⇔ sign bit set or exponent all ones ⇔ high word ≥ 0x7ff00000 as unsigned (the FP64 form
    // `!(v >= 0 && v < Inf)` compiled to ≈ 14 integer instructions per value) — and the largest high word decides for
    // all of them: a running unsigned maximum (three-input VIMNMX3: one instruction per two values) and ONE compare.
    // −0.0 is flagged too; its group then runs with t / p = 1 and writes +0.0, as the reference's `ifelse(…, 0)` does.
    unsigned highest = 0u;
    for (int t = 0; t < a.ntracers; t++) {
        if (STAGED) {
            highest = max(highest, reinterpret_cast<const unsigned*>(mine + t * SN_BLOCK)[1]);  // the high word alone
        } else {
            const double v = a.tracers[t][idx];
            mine[t * SN_BLOCK] = v;
            highest = max(highest, (unsigned)__double2hiint(v));
        }
    }
    const bool touched = highest >= 0x7ff00000u;
    if (!__any_sync(__activemask(), touched)) return;  // warp-uniform: the common case reads its cells and leaves
    unsigned dirty = 0;  // bit t ⇔ tracer t was rescaled and must be written back
    for (int q = 0; q < a.ngroups; q++) {
        const obm_scale_group& g = a.groups[q];
        double t = 0.0, p = 0.0;
        unsigned members = 0;
        bool bad = false;  // some member is negative (−0.0 included) or non-finite
        for (int m = 0; m < g.n; m++) {  // negative_tracers.jl:256-264 (same operation order, no FMA contraction)
            const double v = mine[g.index[m] * SN_BLOCK];
            const double s = __dmul_rn(v, g.scalefactor[m]);
            t = __dadd_rn(t, s);
            if (v > 0) p = __dadd_rn(p, s);
            bad |= (unsigned)__double2hiint(v) >= 0x7ff00000u;
            members |= 1u << g.index[m];
        }
        // No negative and no non-finite member: the reference would multiply every member by t/p = 1 (up to 1 ulp of
        // (v·t)/t rounding noise) — leave the cell alone.  The decision is taken from the members themselves, not from
        // p == t: a negative member far below the ulp of the positive sum (DIC = 2000, P = −1e-14) is absorbed by the
        // rounding of t, yet the reference still zeroes it (:268-274).
        if (!bad) continue;
        t = t < 0 ? a.fill : t;             // :266
        const double ratio = __ddiv_rn(t, p);  // one division per group; v·t/p ≡ v·(t/p) to ≤ 1 ulp
        for (int m = 0; m < g.n; m++) {     // :268-274
            const double v = mine[g.index[m] * SN_BLOCK];
            const bool keep = !isfinite(v) | (v > 0);
            mine[g.index[m] * SN_BLOCK] = keep ? __dmul_rn(v, ratio) : 0.0;
        }
        dirty |= members;
    }
    if (dirty == 0) return;
    for (int t = 0; t < a.ntracers; t++)
        if ((dirty >> t) & 1u) a.tracers[t][idx] = mine[t * SN_BLOCK];
}

__global__ void __launch_bounds__(SN_BLOCK) scale_negative_kernel(const __grid_constant__ ScaleArgs a) {
    extern __shared__ double sm[];  // [ntracers][SN_BLOCK]
    int i, j, k;
    if (!thread_cell(a.d, i, j, k)) return;
    if (immersed_cell(a.d, i, j, k)) return;  // negative_tracers.jl:194,253
    scale_cell(a, sm + threadIdx.x, cell_index(a.d, i, j, k));
}

// ---- negative scaling + calcite saturation of the same cell in one pass (PISCES stage prologue) ----------------------
// `update_biogeochemical_state!` runs the modifiers (OceanBioME.jl:161-169) and, for PISCES, later
// `compute_calcite_saturation!` (PISCES/update_state.jl:13) on the rescaled DIC, Alk, Si.  Both are pointwise, nothing
// in between writes those tracers, T or S, so the Ω solve can consume the staged, already rescaled values of the cell:
// the HBM-bound scaling pass (8 B × distinct tracers per cell) disappears under the FP64-bound carbonate solve, and
// DIC, Alk, Si are read once instead of twice.  Same device functions as the separate kernels ⇒ same results.
struct ScaleOmegaArgs {
    ScaleArgs s;
    const double *T, *S, *DIC, *Alk, *Si;  // used directly when the field is not one of s.tracers
    int iDIC, iAlk, iSi;                  // index in s.tracers, or −1
    double *Omega, *Hst;
    int iterations;
    double H_init;
    const cc::LevelTables* levels;  // OBM_SN_LEVEL == 2: per-level tables of this launch's grid in global memory, or nullptr
};

#ifndef OBM_SN_ASYNC
#define OBM_SN_ASYNC 1
#endif
__device__ __forceinline__ void cp_async8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
#ifndef OBM_SN_MIN_BLOCKS
#define OBM_SN_MIN_BLOCKS 8  // 64 registers + 96 B of spill; 7: 72, none; 6: 80, none — timed in profiles/r03_kernel_variants.txt (8 blocks: −9 %)
#endif
// OBM_SN_LEVEL: the per-level TEOS-10 / pressure-correction tables of carbon_chemistry.cuh in this kernel.  Timed on the
// B200 (profiles/r03_kernel_variants.txt): with 8 resident blocks the tables' block barrier costs more than the ≈ 60 FP64
// instructions per cell they save (1.026 ms with, 0.976 ms without, per 16.8 M cells), so they are off here; the
// stand-alone Ω kernel keeps them (OBM_CC_LEVEL).
#ifndef OBM_SN_LEVEL
#define OBM_SN_LEVEL 0
#endif
// Per-level tables in global memory (OBM_SN_LEVEL == 2).  Static storage — the C ABI hands the library no workspace — as a
// ring: every launch takes the next slot, so launches in flight on different streams with different grids do not share
// one (up to LEVEL_RING of them; launches on the same grid write identical values anyway).
constexpr int LEVEL_RING = 8;
constexpr int LEVEL_MAX = 512;  // deeper grids use the direct form
__device__ cc::LevelTables g_level_tables[LEVEL_RING][LEVEL_MAX];
__global__ void __launch_bounds__(64) level_tables_kernel(cc::LevelTables* out, const double* zc) {
    // the same pressure expression as the cells of this level evaluate (compute_calcite_saturation.jl:27)
    cc::fill_level_entry(out[blockIdx.x], fabs(zc[blockIdx.x]) * 9.80665 * 1026.0 / 100000.0, threadIdx.x);
}

__global__ void __launch_bounds__(SN_BLOCK, OBM_SN_MIN_BLOCKS) scale_negative_calcite_kernel(const __grid_constant__ ScaleOmegaArgs a) {
    extern __shared__ double sm[];  // [ntracers][SN_BLOCK]
    int i = 0, j = 0, k = 0;
    const bool inside = thread_cell(a.s.d, i, j, k);
    const long long idx = cell_index(a.s.d, i, j, k);
    double* mine = sm + threadIdx.x;
    // an immersed cell is not rescaled (negative_tracers.jl:194,253); Ω is still computed there, from the values as they
    // are — compute_calcite_saturation! has no such guard (PISCES/compute_calcite_saturation.jl:9-37)
    const bool dry = inside && immersed_cell(a.s.d, i, j, k);
#if OBM_SN_ASYNC
    // The cell's tracers travel HBM → shared memory by cp.async (no registers, nothing waits on them) while the thread
    // evaluates everything of the Ω solve that needs T, S and the pressure only — density, seven equilibrium constants,
    // the solubility product: half of the kernel's arithmetic.  ncu (r3b) had 19 % of this kernel's stall samples on the
    // first use of the tracer loads, which used to come first.  A thread reads back only what it copied itself: no barrier.
    if (inside && !dry)
        for (int t = 0; t < a.s.ntracers; t++) cp_async8(mine + t * SN_BLOCK, a.s.tracers[t] + idx);
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
#if OBM_SN_LEVEL == 2
    // the block's z-level (blockIdx.z) fixes the pressure: its TEOS-10 and pressure-correction tables come from a table
    // of all levels that a 3 µs launch built just before this one (level_tables_kernel) — block-uniform addresses, read
    // through L1 like constants, no shared memory and no barrier
    const cc::LevelTables* lvl = a.levels ? a.levels + blockIdx.z : nullptr;
#elif OBM_SN_LEVEL
    // the block's z-level (blockIdx.z) fixes the pressure: its TEOS-10 and pressure-correction tables, built once
    __shared__ cc::LevelTables level;
    cc::fill_level_entry(level, fabs(a.s.d.zc[blockIdx.z]) * 9.80665 * 1026.0 / 100000.0, threadIdx.x);
    __syncthreads();
    const cc::LevelTables* lvl = &level;
#else
    const cc::LevelTables* lvl = nullptr;
#endif
    if (!inside) return;
    const double P = fabs(a.s.d.zc[k]) * 9.80665 * 1026.0 / 100000.0;  // compute_calcite_saturation.jl:27
    const double T = a.T[idx], S = a.S[idx];  // never rescaled (not members of any conserved group)
#if OBM_SN_ASYNC
    cc::Prepared q;
    cc::prepare<true>(q, true, T, S, P, true, false, lvl, true);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!dry) scale_cell<true>(a.s, mine, idx);
#else
    if (!dry) scale_cell(a.s, mine, idx);
#endif
    const double DIC = (a.iDIC >= 0 && !dry) ? mine[a.iDIC * SN_BLOCK] : a.DIC[idx];
    const double Alk = (a.iAlk >= 0 && !dry) ? mine[a.iAlk * SN_BLOCK] : a.Alk[idx];
    const double Si = (a.iSi >= 0 && !dry) ? mine[a.iSi * SN_BLOCK] : a.Si[idx];
#if OBM_SN_ASYNC
    a.Omega[idx] = cc::finish<true>(q, true, OBM_CC_OMEGA_CALCITE, S, DIC, Alk, P, true, Si, false, 0.0, false, 0.0, a.H_init,
                                    a.iterations, a.Hst ? a.Hst + idx : nullptr, lvl);
#else
    a.Omega[idx] = cc::solve<true>(OBM_CC_OMEGA_CALCITE, T, S, DIC, Alk, P, true, Si, false, 0.0, false, 0.0, a.H_init,
                                   a.iterations, a.Hst ? a.Hst + idx : nullptr, lvl);
#endif
}

struct ZeroArgs {
    long long n;
    int ntracers;
    double* tracers[OBM_MAX_SCALE_TRACERS];
};
__global__ void __launch_bounds__(256) zero_negative_kernel(const __grid_constant__ ZeroArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int t = 0; t < a.ntracers; t++) {
        double* c = a.tracers[t];
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < a.n; q += stride) {
            const double v = c[q];
            c[q] = jl_max(0.0, v);  // parent .= max.(0.0, parent), NaN-propagating, -0.0 → 0.0
        }
    }
}

// ---- inventory: out[g] = Σ_cells (Σ_f sf·c_f)·V — deterministic two-level tree, no double atomics ----
// ONE pass over the cells for all groups: every distinct tracer is read once per cell (8 B × ntracers, HBM-bound), the
// groups are rows of a dense weight table W[g][t] (0 where tracer t is not a member of group g) held in the constant
// bank.  Blocks walk (j, k) rows, threads stride along x: coalesced, one integer division per ROW.
// OBM_INV_BATCH tracers' loads issued before the first is used.  Timed at 33.5 M cells (scripts/time_inventory.py, r5f): 1 → 1.48 ms
// (3.6 TB/s), 4 → 2.11 ms, 8 → 2.59 ms — the registers the batch holds cost more warps than the loads in flight gain; 8 blocks per SM: ± 0.
#ifndef OBM_INV_BATCH
#define OBM_INV_BATCH 1
#endif
// Blocks per SM of the persistent grid.  With the kernel templated on the number of groups PISCES' five budgets take 46 registers,
// so five blocks of 256 threads are resident: 33.5 M cells in 1.48 ms (r03: 16 accumulators, 61 registers, 4 blocks) → 1.06 ms (4 blocks)
// → 0.94 ms (5 blocks) = 5.7 TB/s, 0.87 of the measured HBM peak; 6 → 1.38 ms (a tail wave), 8 → 1.09 ms (scripts/time_inventory.py, r5g).
#ifndef OBM_INV_BLOCKS_PER_SM
#define OBM_INV_BLOCKS_PER_SM 5
#endif
constexpr int INV_BLOCKS = 148 * OBM_INV_BLOCKS_PER_SM;
constexpr int INV_THREADS = 256;

struct InvArgs {
    GridDims d;
    int ntracers, ngroups;
    const double* tracers[OBM_MAX_SCALE_TRACERS];
    double w[OBM_MAX_SCALE_GROUPS][OBM_MAX_SCALE_TRACERS];
    const double* volume;
    double uniform_volume;
    double* partial;  // [ngroups][INV_BLOCKS]
    double* out;
};

// NG = the number of groups, a compile-time constant: acc / s cost 2·NG registers instead of 16 (PISCES' five budgets: 61 → ≈ 45
// registers), and every resident warp more is another row of loads in flight — what this kernel lives on (see OBM_INV_BATCH)
template <int NG>
__global__ void __launch_bounds__(INV_THREADS) inventory_partial_kernel(const __grid_constant__ InvArgs a) {
    __shared__ double red[NG][INV_THREADS / 32];
    const GridDims& d = a.d;
    const int nx = d.i1 - d.i0, ny = d.j1 - d.j0;
    const int rows = ny * d.Nz;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) acc[g] = 0.0;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int k = row / ny;
        const int j = d.j0 + (row - k * ny);
        for (int ii = threadIdx.x; ii < nx; ii += INV_THREADS) {
            const long long idx = cell_index(d, d.i0 + ii, j, k);
            double s[NG];
#pragma unroll
            for (int g = 0; g < NG; g++) s[g] = 0.0;
            // (OBM_INV_BATCH > 1: that many tracers' loads in flight before the first is used — measured slower, see above)
            int t = 0;
            for (; t + OBM_INV_BATCH <= a.ntracers; t += OBM_INV_BATCH) {
                double v[OBM_INV_BATCH];
#pragma unroll
                for (int q = 0; q < OBM_INV_BATCH; q++) v[q] = __ldcs(a.tracers[t + q] + idx);  // streamed: read once
#pragma unroll
                for (int q = 0; q < OBM_INV_BATCH; q++)
#pragma unroll
                    for (int g = 0; g < NG; g++)
                        s[g] = fma(a.w[g][t + q], v[q], s[g]);
            }
            for (; t < a.ntracers; t++) {
                const double v = __ldcs(a.tracers[t] + idx);
#pragma unroll
                for (int g = 0; g < NG; g++)
                    s[g] = fma(a.w[g][t], v, s[g]);
            }
            const double V = a.volume ? a.volume[idx] : a.uniform_volume;
#pragma unroll
            for (int g = 0; g < NG; g++) acc[g] = fma(s[g], V, acc[g]);
        }
    }
#pragma unroll
    for (int g = 0; g < NG; g++) {
        double v = acc[g];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[g][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int g = 0; g < NG; g++) {
            double v = lane < INV_THREADS / 32 ? red[g][lane] : 0.0;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) a.partial[g * INV_BLOCKS + blockIdx.x] = v;
        }
    }
}

__global__ void __launch_bounds__(INV_THREADS) inventory_final_kernel(const double* partial, int nblocks, double* out) {
    __shared__ double red[INV_THREADS / 32];
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) acc += partial[q * nblocks + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        double v = lane < INV_THREADS / 32 ? red[lane] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) out[q] = v;
    }
}

static int check_groups(const char* who, int ntracers, int ngroups, const obm_scale_group* groups) {
    OBM_REQUIRE(ntracers >= 1 && ntracers <= OBM_MAX_SCALE_TRACERS, OBM_ESIZE, "%s: ntracers = %d outside [1, %d]", who,
                ntracers, OBM_MAX_SCALE_TRACERS);
    OBM_REQUIRE(ngroups >= 0 && ngroups <= OBM_MAX_SCALE_GROUPS, OBM_ESIZE, "%s: ngroups = %d outside [0, %d]", who, ngroups,
                OBM_MAX_SCALE_GROUPS);
    OBM_REQUIRE(ngroups == 0 || groups != nullptr, OBM_ENULL, "%s: groups is NULL", who);
    for (int q = 0; q < ngroups; q++) {
        OBM_REQUIRE(groups[q].n >= 1 && groups[q].n <= OBM_MAX_GROUP_SIZE, OBM_ESIZE, "%s: group %d has n = %d", who, q,
                    groups[q].n);
        for (int m = 0; m < groups[q].n; m++)
            OBM_REQUIRE(groups[q].index[m] >= 0 && groups[q].index[m] < ntracers, OBM_ESIZE,
                        "%s: group %d member %d indexes tracer %d of %d", who, q, m, groups[q].index[m], ntracers);
    }
    return 0;
}

}  // namespace obm

using namespace obm;

extern "C" int obm_scale_negative_tracers(const obm_grid* grid, int ntracers, double* const* tracers, int ngroups,
                                          const obm_scale_group* groups, double invalid_fill_value, void* stream) {
    OBM_REQUIRE(tracers != nullptr, OBM_ENULL, "obm_scale_negative_tracers: tracers is NULL");
    int rc = check_groups("obm_scale_negative_tracers", ntracers, ngroups, groups);
    if (rc) return rc;
    if (ngroups == 0) return 0;
    ScaleArgs a;
    memset(&a, 0, sizeof(a));
    rc = make_dims(grid, &a.d, false);
    if (rc) return rc;
    a.ntracers = ntracers;
    a.ngroups = ngroups;
    a.fill = invalid_fill_value;
    for (int t = 0; t < ntracers; t++) {
        OBM_REQUIRE(tracers[t] != nullptr, OBM_ENULL, "obm_scale_negative_tracers: tracers[%d] is NULL", t);
        a.tracers[t] = tracers[t];
    }
    for (int q = 0; q < ngroups; q++) a.groups[q] = groups[q];
    const size_t smem = (size_t)ntracers * SN_BLOCK * sizeof(double);
    scale_negative_kernel<<<cell_grid(a.d, SN_BLOCK), SN_BLOCK, smem, (cudaStream_t)stream>>>(a);
    return launch_status("scale_negative_kernel");
}

extern "C" int obm_scale_negative_tracers_calcite_saturation(const obm_grid* grid, int ntracers, double* const* tracers,
                                                             int ngroups, const obm_scale_group* groups,
                                                             double invalid_fill_value, const obm_carbchem_params* p,
                                                             const double* T, const double* S, const double* DIC,
                                                             const double* Alk, const double* Si, double* Omega,
                                                             double* H_state, void* stream) {
    const char* who = "obm_scale_negative_tracers_calcite_saturation";
    OBM_REQUIRE(tracers != nullptr, OBM_ENULL, "%s: tracers is NULL", who);
    OBM_REQUIRE(T && S && DIC && Alk && Si && Omega, OBM_ENULL, "%s: a field pointer is NULL", who);
    int rc = check_groups(who, ntracers, ngroups, groups);
    if (rc) return rc;
    static thread_local ScaleOmegaArgs a;
    memset(&a, 0, sizeof(a));
    rc = make_dims(grid, &a.s.d, true);
    if (rc) return rc;
    a.s.ntracers = ntracers;
    a.s.ngroups = ngroups;
    a.s.fill = invalid_fill_value;
    a.iDIC = a.iAlk = a.iSi = -1;
    for (int t = 0; t < ntracers; t++) {
        OBM_REQUIRE(tracers[t] != nullptr, OBM_ENULL, "%s: tracers[%d] is NULL", who, t);
        OBM_REQUIRE(tracers[t] != T && tracers[t] != S, OBM_ESIZE, "%s: T and S may not be rescaled tracers", who);
        a.s.tracers[t] = tracers[t];
        if (tracers[t] == DIC) a.iDIC = t;
        if (tracers[t] == Alk) a.iAlk = t;
        if (tracers[t] == Si) a.iSi = t;
    }
    for (int q = 0; q < ngroups; q++) a.s.groups[q] = groups[q];
    a.T = T; a.S = S; a.DIC = DIC; a.Alk = Alk; a.Si = Si; a.Omega = Omega; a.Hst = H_state;
    a.iterations = (p && p->newton_iterations > 0) ? p->newton_iterations : 12;
    a.H_init = pow(10.0, -((p && p->initial_pH_guess > 0) ? p->initial_pH_guess : 8.0));
    const size_t smem = (size_t)ntracers * SN_BLOCK * sizeof(double);
    a.levels = nullptr;
#if OBM_SN_LEVEL == 2
    if (a.s.d.Nz <= LEVEL_MAX) {
        static cc::LevelTables* base = [] {
            void* q = nullptr;
            return cudaGetSymbolAddress(&q, g_level_tables) == cudaSuccess ? (cc::LevelTables*)q : nullptr;
        }();
        static std::atomic<unsigned> next{0};
        if (base) {
            cc::LevelTables* slot = base + (size_t)(next.fetch_add(1u) % LEVEL_RING) * LEVEL_MAX;
            level_tables_kernel<<<a.s.d.Nz, 64, 0, (cudaStream_t)stream>>>(slot, a.s.d.zc);
            rc = launch_status("level_tables_kernel");
            if (rc) return rc;
            a.levels = slot;
        }
    }
#endif
    scale_negative_calcite_kernel<<<cell_grid(a.s.d, SN_BLOCK), SN_BLOCK, smem, (cudaStream_t)stream>>>(a);
    return launch_status("scale_negative_calcite_kernel");
}

extern "C" int obm_zero_negative_tracers(int64_t n_parent, int ntracers, double* const* tracers, void* stream) {
    OBM_REQUIRE(tracers != nullptr, OBM_ENULL, "obm_zero_negative_tracers: tracers is NULL");
    OBM_REQUIRE(ntracers >= 0 && ntracers <= OBM_MAX_SCALE_TRACERS && n_parent >= 0, OBM_ESIZE,
                "obm_zero_negative_tracers: ntracers = %d, n = %lld", ntracers, (long long)n_parent);
    if (ntracers == 0 || n_parent == 0) return 0;
    ZeroArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n_parent;
    a.ntracers = ntracers;
    for (int t = 0; t < ntracers; t++) {
        OBM_REQUIRE(tracers[t] != nullptr, OBM_ENULL, "obm_zero_negative_tracers: tracers[%d] is NULL", t);
        a.tracers[t] = tracers[t];
    }
    long long blocks = (n_parent + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    zero_negative_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_status("zero_negative_kernel");
}

extern "C" int64_t obm_inventory_workspace_bytes(int ngroups) {
    if (ngroups < 0 || ngroups > OBM_MAX_SCALE_GROUPS) return OBM_ESIZE;
    return (int64_t)ngroups * INV_BLOCKS * (int64_t)sizeof(double);
}

extern "C" int obm_inventory(const obm_grid* grid, int ntracers, const double* const* tracers, int ngroups,
                             const obm_scale_group* groups, const double* cell_volume, double uniform_volume, double* out,
                             void* workspace, void* stream) {
    OBM_REQUIRE(tracers && out && workspace, OBM_ENULL, "obm_inventory: tracers / out / workspace is NULL");
    int rc = check_groups("obm_inventory", ntracers, ngroups, groups);
    if (rc) return rc;
    if (ngroups == 0) return 0;
    static thread_local InvArgs a;  // 2.4 KB of kernel arguments, kept off the caller's stack
    memset(&a, 0, sizeof(a));
    rc = make_dims(grid, &a.d, false);
    if (rc) return rc;
    a.ntracers = ntracers;
    a.ngroups = ngroups;
    for (int t = 0; t < ntracers; t++) {
        OBM_REQUIRE(tracers[t] != nullptr, OBM_ENULL, "obm_inventory: tracers[%d] is NULL", t);
        a.tracers[t] = tracers[t];
    }
    for (int q = 0; q < ngroups; q++)
        for (int m = 0; m < groups[q].n; m++) a.w[q][groups[q].index[m]] += groups[q].scalefactor[m];
    a.volume = cell_volume;
    a.uniform_volume = uniform_volume;
    a.partial = (double*)workspace;
    a.out = out;
    cudaStream_t s = (cudaStream_t)stream;
    switch (ngroups) {
        case 1: inventory_partial_kernel<1><<<INV_BLOCKS, INV_THREADS, 0, s>>>(a); break;
        case 2: inventory_partial_kernel<2><<<INV_BLOCKS, INV_THREADS, 0, s>>>(a); break;
        case 3: inventory_partial_kernel<3><<<INV_BLOCKS, INV_THREADS, 0, s>>>(a); break;
        case 4: inventory_partial_kernel<4><<<INV_BLOCKS, INV_THREADS, 0, s>>>(a); break;
        case 5: inventory_partial_kernel<5><<<INV_BLOCKS, INV_THREADS, 0, s>>>(a); break;
        case 6: inventory_partial_kernel<6><<<INV_BLOCKS, INV_THREADS, 0, s>>>(a); break;
        case 7: inventory_partial_kernel<7><<<INV_BLOCKS, INV_THREADS, 0, s>>>(a); break;
        default: inventory_partial_kernel<8><<<INV_BLOCKS, INV_THREADS, 0, s>>>(a); break;
    }
    rc = launch_status("inventory_partial_kernel");
    if (rc) return rc;
    inventory_final_kernel<<<ngroups, INV_THREADS, 0, s>>>(a.partial, INV_BLOCKS, out);
    return launch_status("inventory_final_kernel");
}
