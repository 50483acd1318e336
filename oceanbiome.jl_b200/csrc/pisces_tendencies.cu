// pisces_tendencies.cu — ONE fused kernel for the 24 PISCES tracer tendencies.
//
// Replaces the 24 per-tracer callables of src/Models/AdvectedPopulations/PISCES/ which
// Oceananigans evaluates in 26 compute_Gc! launches per stage, each re-reading its
// inputs and re-evaluating nutrient limitation / growth rates / grazing / mortalities / bacteria /
// aggregation / iron chemistry (≈ 24 × 15 field reads and ≥ 10× redundant FP64 work per cell).
// Here a thread owns a cell: 23 tracers (DIC, Alk, S are never read) + PAR₁₂₃, PAR, Ω, w (two
// faces of two fields) are read once, coalesced along x; every shared sub-model is evaluated once
// in registers in the reference's operation order; temperature powers b^T are evaluated once per
// DISTINCT base; each tendency is stored as soon as its last term is known.
//
// FP64-pipe bound: ≈ 10 exp-class calls + 1 pow + 1 sqrt + ≈ 60 divisions per cell (vs ≈ 110
// divisions and ≈ 90 transcendental calls of the un-fused reference path).
// Reference quirks are reproduced on purpose: swapped day-length arguments
// (host-evaluated, obm_pisces_params), (P, D, POC, Z) prey order, negative flux-feeding "flux",
// Alk = NH₄ − NO₃ − 2·CaCO₃.
#include <stdlib.h>

#include "pisces_cell.cuh"

namespace obm {

__device__ __forceinline__ unsigned nonfinite(double t) {
    return ((unsigned)(__double2hiint(t)) & 0x7ff00000u) == 0x7ff00000u;  // exponent test on the high word
}
__device__ __forceinline__ void red_add(double* p, double t) {  // fire-and-forget Gⁿ += t at the L2
    atomicAdd(p, t);  // result unused ⇒ RED.E.ADD.F64
}

// Where a tendency goes once its last term is known.
// FAST pass: a finite result is final — stored (or added, accumulate mode) to Gⁿ at once, so it leaves the register
// file; a non-finite one is only marked.  EXACT pass (rare): recomputes the cell with the reference's operation
// semantics and delivers exactly the marked tendencies.  Every tendency is delivered once.
template <bool ACC, bool FULL>  // ACC: Gⁿ += t; FULL: every tendency has a destination (no per-tendency mask test)
struct FastSink {
    const PiscesArgs& a;
    long long idx;
    unsigned pending;
    unsigned hold;  // 1: deliver nothing, mark everything (a cell whose NaN inputs min/max would swallow)
    __device__ __forceinline__ void put(int n, double t) {
        if (!FULL && !((a.out_mask >> n) & 1u)) return;
        const unsigned nf = nonfinite(t) | hold;
        pending |= nf << n;
        if (!nf) {
            if (ACC) red_add(a.g[n] + idx, t);
            else a.g[n][idx] = t;
        }
    }
};
struct ExactSink {
    const PiscesArgs& a;
    long long idx;
    unsigned pending;
    __device__ __forceinline__ void put(int n, double t) {
        if (!((pending >> n) & 1u)) return;
        if (a.accumulate) a.g[n][idx] += t;
        else a.g[n][idx] = t;
    }
};

// A NaN that only flows through min/max would be swallowed by the FAST pass's selects — an input NaN directly, an
// input ±Inf through Inf − Inf or Inf / Inf on the way: a cell with ANY non-finite input goes straight to the EXACT
// pass.  Integer test on the high words (exponent all ones), min-reduced: ≈ 70 ALU-pipe instructions per cell.
__device__ __forceinline__ bool needs_exact(const Inputs& in) {
    static_assert(sizeof(Inputs) % sizeof(double) == 0, "Inputs is all doubles");
    const double* v = reinterpret_cast<const double*>(&in);
    unsigned gap = 0x7ff00000u;
#pragma unroll
    for (int n = 0; n < (int)(sizeof(Inputs) / sizeof(double)); n++)
        gap = min(gap, ~(unsigned)__double2hiint(v[n]) & 0x7ff00000u);  // 0 ⇔ NaN or ±Inf
    return gap == 0u;
}


__device__ __forceinline__ Inputs load_inputs(const PiscesArgs& a, long long idx, long long pl, int k) {
    // 3-D inputs are read exactly once: keep them out of L1 (ld.global.cg), which is left to the spill slots
#ifndef OBM_PISCES_LD
#define OBM_PISCES_LD __ldcg
#endif
    auto ld = [](const double* p) { return OBM_PISCES_LD(p); };
    Inputs in;
    in.P = ld(a.c[T_P] + idx); in.PChl = ld(a.c[T_PChl] + idx); in.PFe = ld(a.c[T_PFe] + idx);
    in.D = ld(a.c[T_D] + idx); in.DChl = ld(a.c[T_DChl] + idx); in.DFe = ld(a.c[T_DFe] + idx); in.DSi = ld(a.c[T_DSi] + idx);
    in.Z = ld(a.c[T_Z] + idx); in.M = ld(a.c[T_M] + idx); in.DOC = ld(a.c[T_DOC] + idx);
    in.POC = ld(a.c[T_POC] + idx); in.GOC = ld(a.c[T_GOC] + idx); in.SFe = ld(a.c[T_SFe] + idx); in.BFe = ld(a.c[T_BFe] + idx);
    in.PSi = ld(a.c[T_PSi] + idx); in.CaCO3 = ld(a.c[T_CaCO3] + idx);
    in.c.NO3 = ld(a.c[T_NO3] + idx); in.c.NH4 = ld(a.c[T_NH4] + idx); in.c.PO4 = ld(a.c[T_PO4] + idx); in.c.Fe = ld(a.c[T_Fe] + idx);
    in.c.Si = ld(a.c[T_Si] + idx); in.c.O2 = ld(a.c[T_O2] + idx); in.c.T = ld(a.c[T_T] + idx);
    in.c.PAR1 = ld(a.f.PAR1 + idx); in.c.PAR2 = ld(a.f.PAR2 + idx); in.c.PAR3 = ld(a.f.PAR3 + idx);
    in.PARt = ld(a.f.PAR + idx); in.Omega = ld(a.f.Omega + idx);
    // ℑzᵃᵃᶜ(i, j, k, grid, w) = (w[k] + w[k+1]) / 2 — two_size_class.jl:95-98
    in.wPOC = (ld(a.f.wPOC + idx) + ld(a.f.wPOC + idx + a.d.sz)) / 2;
    in.wGOC = (ld(a.f.wGOC + idx) + ld(a.f.wGOC + idx + a.d.sz)) / 2;
    in.c.zmxl = a.f.mixed_layer_depth_xy[pl];
    in.c.zeu = a.f.euphotic_depth_xy[pl];
    in.c.kappa = a.f.mean_mixed_layer_vertical_diffusivity_xy[pl];
    in.mlPAR = a.f.mean_mixed_layer_light_xy[pl];
    in.c.z = a.d.zc[k];
    return in;
}

// The rare path (non-finite results): re-reads the cell and evaluates it with the reference's exact
// operation semantics.  Kept out of line so that it does not weigh on the fast path's registers.
__device__ __noinline__ void cell_exact(const PiscesArgs& a, long long idx, long long pl, int k, unsigned pending) {
    const Inputs in = load_inputs(a, idx, pl, k);
    ExactSink sink{a, idx, pending};
    cell_tendencies<true>(a, in, sink);
}

// 3 blocks of 128 threads per SM (168 registers, ≈ 116 B of spill per thread): the spill slots of all resident
// threads then stay inside L1, which the input loads bypass (ld.global.cg).  4 blocks (128 registers, ≈ 410 B of
// spill) is 2 % slower, 5 blocks thrashes L1.
#ifndef OBM_PISCES_MIN_BLOCKS
#define OBM_PISCES_MIN_BLOCKS 3
#endif
#ifdef OBM_PISCES_MAXNREG
#define OBM_PISCES_BOUNDS __maxnreg__(OBM_PISCES_MAXNREG)
#else
#define OBM_PISCES_BOUNDS __launch_bounds__(PB, OBM_PISCES_MIN_BLOCKS)
#endif

// One thread per cell.  Measured on B200 (16.8 M cells, accumulate mode): 2.75 ms whether the block is 64…512
// threads, whether 3 or 4 blocks are resident (168 / 128 registers), with or without shared-memory staging of the
// inputs or of Gⁿ, L2 prefetch of the next wave, or block-wide lock-step — see DESIGN.md "PISCES kernel: what bounds it".
template <bool ACC, bool FULL>
__global__ void OBM_PISCES_BOUNDS pisces_tendency_kernel(const __grid_constant__ PiscesArgs a) {
    int i, j, k;
    if (!thread_cell(a.d, i, j, k)) return;
    const long long idx = cell_index(a.d, i, j, k);
    const long long pl = plane_index(a.d, i, j);

    // ---- one coalesced read of the cell ------------------------------------------------------------
    const Inputs in = load_inputs(a, idx, pl, k);

    // The NaN-input guard is a data dependency of the stores, not a branch ahead of the arithmetic: a branch on
    // loaded values here made the compiler sink two thirds of the input loads below it, i.e. two DRAM round trips
    // per cell instead of one (profiles/r02_pisces_tendency_hot_lines.txt: 24 % of the samples on three waits).
    FastSink<ACC, FULL> sink{a, idx, 0u, needs_exact(in) ? 1u : 0u};
    cell_tendencies<false>(a, in, sink);
    if (sink.pending) cell_exact(a, idx, pl, k, sink.pending);  // rare: non-finite results, NaN inputs
}

// ---- ModelLatitude (PISCES/common.jl:27-28: φ = φnode(i, j, k, grid)) ------------------------------------------------------------
// On a grid that carries its own latitude the three host-evaluated quantities of obm_pisces_params that depend on it —
// latitude, day_length(φ, t) with the reference's swapped arguments (growth_rate.jl:29-30) and day_length(t, φ)
// (:141-143) — differ from row to row.  A block works on one (j, k) row (thread_cell), so it copies the kernel
// arguments to shared memory, patches those three values and the parameter-only sub-expressions derived from them
// (pisces_prepare's formulas, IEEE arithmetic: bit-identical to the host evaluation) for ITS row, and runs the very same
// cell code on the patched copy.  Nothing changes in the prescribed-latitude kernel above, which every named
// configuration uses; this variant reads its parameters from shared memory instead of the constant bank.
template <bool ACC>
__global__ void OBM_PISCES_BOUNDS pisces_tendency_rows_kernel(const __grid_constant__ PiscesArgs a, const double* __restrict__ rows) {
    __shared__ PiscesArgs sa;
    static_assert(sizeof(PiscesArgs) % sizeof(double) == 0, "PiscesArgs is copied in 8-byte words");
    {
        const double* src = reinterpret_cast<const double*>(&a);
        double* dst = reinterpret_cast<double*>(&sa);
        for (int w = threadIdx.x; w < (int)(sizeof(PiscesArgs) / sizeof(double)); w += blockDim.x) dst[w] = src[w];
    }
    __syncthreads();
    int i = 0, j = 0, k = 0;
    const bool inside = thread_cell(a.d, i, j, k);
    if (threadIdx.x == 0) {  // always inside: its block exists
        const int Ny = a.d.Ny;
        const double lat = rows[j], dlg = rows[Ny + j], dlc = rows[2 * Ny + j];
        sa.p.latitude = lat;
        sa.p.day_length_growth = dlg;
        sa.p.day_length_chlorophyll = dlc;
        sa.dv.f1_growth = 1.5 * dlg / (dlg + 0.5 * DAY);
        sa.dv.dl_over_f1_chl = dlc / (1.5 * dlc / (dlc + 0.5 * DAY));
        sa.dv.inv_resp[0] = 1.0 / (dlg * (a.p.nano.basal_respiration_rate + a.p.nano.reference_growth_rate));
        sa.dv.inv_resp[1] = 1.0 / (dlg * (a.p.diatoms.basal_respiration_rate + a.p.diatoms.reference_growth_rate));
    }
    __syncthreads();
    if (!inside) return;
    const long long idx = cell_index(a.d, i, j, k);
    const long long pl = plane_index(a.d, i, j);
    const Inputs in = load_inputs(a, idx, pl, k);
    FastSink<ACC, false> sink{sa, idx, 0u, needs_exact(in) ? 1u : 0u};
    cell_tendencies<false>(sa, in, sink);
    if (sink.pending) cell_exact(sa, idx, pl, k, sink.pending);
}

// ---- the same cell arithmetic, software-pipelined over a persistent launch (build option OBM_PISCES_PIPE) -------------------------
// ncu (r3a, source page): ≈ 15 % of the warp-state samples of the kernel above sit on the first use of a loaded input — a
// thread's 38 loads go out together and, at 12 warps per SM, little else is ready while they are in flight.  Here a block
// walks a sequence of 128-cell row chunks and every thread copies the inputs of ITS cell of the next chunk into ITS
// shared-memory column (cp.async, 8 bytes each: no registers, no barrier — a thread only reads back what it copied
// itself) while it computes the current one.
#ifndef OBM_PISCES_PIPE
#define OBM_PISCES_PIPE 0
#endif
constexpr int PIPE_ROWS = 36;  // 23 tracers + PAR₁₂₃ + PAR + Ω + 2 × 2 faces of w + 4 column fields
__device__ __forceinline__ void cp_async8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
struct Chunk { int i, j, k; };
__device__ __forceinline__ Chunk chunk_of(const GridDims& d, unsigned t, unsigned chunks, unsigned ny) {
    const unsigned row = t / chunks;
    Chunk r;
    r.i = d.i0 + (int)(t - row * chunks) * PB + (int)threadIdx.x;
    r.k = (int)(row / ny);
    r.j = d.j0 + (int)(row - (unsigned)r.k * ny);
    return r;
}
__device__ __forceinline__ void pipe_issue(const PiscesArgs& a, double* mine, long long idx, long long pl) {
    int r = 0;
#pragma unroll
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++)
        if (!(n == T_DIC || n == T_Alk || n == T_S)) cp_async8(mine + PB * (r++), a.c[n] + idx);
    cp_async8(mine + PB * (r++), a.f.PAR1 + idx); cp_async8(mine + PB * (r++), a.f.PAR2 + idx); cp_async8(mine + PB * (r++), a.f.PAR3 + idx);
    cp_async8(mine + PB * (r++), a.f.PAR + idx); cp_async8(mine + PB * (r++), a.f.Omega + idx);
    cp_async8(mine + PB * (r++), a.f.wPOC + idx); cp_async8(mine + PB * (r++), a.f.wPOC + idx + a.d.sz);
    cp_async8(mine + PB * (r++), a.f.wGOC + idx); cp_async8(mine + PB * (r++), a.f.wGOC + idx + a.d.sz);
    cp_async8(mine + PB * (r++), a.f.mixed_layer_depth_xy + pl); cp_async8(mine + PB * (r++), a.f.euphotic_depth_xy + pl);
    cp_async8(mine + PB * (r++), a.f.mean_mixed_layer_vertical_diffusivity_xy + pl); cp_async8(mine + PB * (r++), a.f.mean_mixed_layer_light_xy + pl);
}
__device__ __forceinline__ Inputs pipe_read(const PiscesArgs& a, const double* mine, int k) {
    Inputs in;
    int r = 0;
    auto ld = [&](int row) { return mine[PB * row]; };
    in.P = ld(r++); in.PChl = ld(r++); in.PFe = ld(r++); in.D = ld(r++); in.DChl = ld(r++);
    in.DFe = ld(r++); in.DSi = ld(r++); in.Z = ld(r++); in.M = ld(r++); in.DOC = ld(r++);
    in.POC = ld(r++); in.GOC = ld(r++); in.SFe = ld(r++); in.BFe = ld(r++); in.PSi = ld(r++);
    in.CaCO3 = ld(r++); in.c.NO3 = ld(r++); in.c.NH4 = ld(r++); in.c.PO4 = ld(r++); in.c.Fe = ld(r++);
    in.c.Si = ld(r++); in.c.O2 = ld(r++); in.c.T = ld(r++);              // tracer order minus DIC, Alk, S
    in.c.PAR1 = ld(r++); in.c.PAR2 = ld(r++); in.c.PAR3 = ld(r++); in.PARt = ld(r++); in.Omega = ld(r++);
    const double wP0 = ld(r++), wP1 = ld(r++), wG0 = ld(r++), wG1 = ld(r++);
    in.wPOC = (wP0 + wP1) / 2;  // ℑzᵃᵃᶜ(i, j, k, grid, w) — two_size_class.jl:95-98
    in.wGOC = (wG0 + wG1) / 2;
    in.c.zmxl = ld(r++); in.c.zeu = ld(r++); in.c.kappa = ld(r++); in.mlPAR = ld(r++);
    in.c.z = a.d.zc[k];
    return in;
}

template <bool ACC, bool FULL>
__global__ void OBM_PISCES_BOUNDS pisces_tendency_pipe_kernel(const __grid_constant__ PiscesArgs a, unsigned ntiles, unsigned chunks) {
    __shared__ double buf[PIPE_ROWS][PB];
    const unsigned ny = (unsigned)(a.d.j1 - a.d.j0);
    unsigned t = blockIdx.x;  // the only loop state that lives across a cell's arithmetic: everything else is recomputed
    {
        const Chunk c = chunk_of(a.d, t, chunks, ny);
        if (c.i < a.d.i1) pipe_issue(a, &buf[0][threadIdx.x], cell_index(a.d, c.i, c.j, c.k), plane_index(a.d, c.i, c.j));
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    do {
        double* mine = &buf[0][threadIdx.x];
        const Chunk c = chunk_of(a.d, t, chunks, ny);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const Inputs in = pipe_read(a, mine, c.k);  // (a thread beyond the row's end reads its idle column: never used)
        t += gridDim.x;
        if (t < ntiles) {  // the column is free (its values are in registers): the next chunk's copies may land
            const Chunk n = chunk_of(a.d, t, chunks, ny);
            if (n.i < a.d.i1) pipe_issue(a, mine, cell_index(a.d, n.i, n.j, n.k), plane_index(a.d, n.i, n.j));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (c.i < a.d.i1) {
            const long long idx = cell_index(a.d, c.i, c.j, c.k);
            FastSink<ACC, FULL> sink{a, idx, 0u, needs_exact(in) ? 1u : 0u};
            cell_tendencies<false>(a, in, sink);
            if (sink.pending) cell_exact(a, idx, plane_index(a.d, c.i, c.j), c.k, sink.pending);
        }
    } while (t < ntiles);
}

}  // namespace obm

using namespace obm;

static int pisces_arguments(PiscesArgs& A, const obm_grid* grid, const obm_pisces_params* p, const double* const* tracers,
                            const obm_pisces_fields* aux, double* const* G, int accumulate) {
    OBM_REQUIRE(p && tracers && aux && G, OBM_ENULL, "obm_pisces_tendencies: params / tracers / aux / G is NULL");
    memset(&A, 0, sizeof(A));
    const int rc = make_dims(grid, &A.d, true);
    if (rc) return rc;
    A.p = *p;
    A.f = *aux;
    OBM_REQUIRE(aux->PAR1 && aux->PAR2 && aux->PAR3 && aux->PAR && aux->Omega && aux->wPOC && aux->wGOC
                    && aux->mixed_layer_depth_xy && aux->euphotic_depth_xy && aux->mean_mixed_layer_vertical_diffusivity_xy
                    && aux->mean_mixed_layer_light_xy,
                OBM_ENULL, "obm_pisces_tendencies: an auxiliary field pointer is NULL");
    OBM_REQUIRE(p->nano.growth_rate_kind >= 0 && p->nano.growth_rate_kind <= 1 && p->diatoms.growth_rate_kind >= 0
                    && p->diatoms.growth_rate_kind <= 1,
                OBM_EENUM, "obm_pisces_tendencies: unknown growth_rate_kind");
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++) {
        const bool read = !(n == T_DIC || n == T_Alk || n == T_S);  // never read by any tendency
        OBM_REQUIRE(!read || tracers[n], OBM_ENULL, "obm_pisces_tendencies: tracers[%d] is NULL", n);
        A.c[n] = tracers[n];
        A.g[n] = (n == T_T || n == T_S) ? nullptr : G[n];  // zero(grid) PISCES.jl:120
    }
    A.accumulate = accumulate ? 1 : 0;
    A.out_mask = 0;
    for (int n = 0; n < NOUT; n++)
        if (A.g[n]) A.out_mask |= 1u << n;
    pisces_prepare(A, p);
    return 0;
}

extern "C" int obm_pisces_tendencies(const obm_grid* grid, const obm_pisces_params* p, const double* const* tracers,
                                     const obm_pisces_fields* aux, double* const* G, int accumulate, void* stream) {
    static thread_local PiscesArgs tl;  // ~2 KB of kernel arguments, kept off the caller's stack
    PiscesArgs& A = tl;
    const int rc = pisces_arguments(A, grid, p, tracers, aux, G, accumulate);
    if (rc) return rc;
    const dim3 gr = cell_grid(A.d, PB);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool carveout_set = [] {  // no shared memory is used: give the whole array to L1 (spill slots)
        cudaFuncSetAttribute(pisces_tendency_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        cudaFuncSetAttribute(pisces_tendency_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        cudaFuncSetAttribute(pisces_tendency_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        cudaFuncSetAttribute(pisces_tendency_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        return true;
    }();
    (void)carveout_set;
    const bool full = A.out_mask == (1u << NOUT) - 1u;
#if OBM_PISCES_PIPE
    {
        const unsigned chunks = (unsigned)((A.d.i1 - A.d.i0 + PB - 1) / PB);
        const long long nt = (long long)chunks * (A.d.j1 - A.d.j0) * A.d.Nz;
        if (nt < (1LL << 31)) {
            static const int sms = [] { int dev = 0, n = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }();
            const long long resident = (long long)sms * OBM_PISCES_MIN_BLOCKS;
            const unsigned blocks = (unsigned)(nt < resident ? nt : resident);
            if (A.accumulate) {
                if (full) pisces_tendency_pipe_kernel<true, true><<<blocks, PB, 0, st>>>(A, (unsigned)nt, chunks);
                else pisces_tendency_pipe_kernel<true, false><<<blocks, PB, 0, st>>>(A, (unsigned)nt, chunks);
            } else {
                if (full) pisces_tendency_pipe_kernel<false, true><<<blocks, PB, 0, st>>>(A, (unsigned)nt, chunks);
                else pisces_tendency_pipe_kernel<false, false><<<blocks, PB, 0, st>>>(A, (unsigned)nt, chunks);
            }
            return launch_status("pisces_tendency_pipe_kernel");
        }
    }
#endif
    if (A.accumulate) {
        if (full) pisces_tendency_kernel<true, true><<<gr, PB, 0, st>>>(A);
        else pisces_tendency_kernel<true, false><<<gr, PB, 0, st>>>(A);
    } else {
        if (full) pisces_tendency_kernel<false, true><<<gr, PB, 0, st>>>(A);
        else pisces_tendency_kernel<false, false><<<gr, PB, 0, st>>>(A);
    }
    return launch_status("pisces_tendency_kernel");
}

extern "C" int obm_pisces_tendencies_rows(const obm_grid* grid, const obm_pisces_params* p, const double* row_latitude_daylengths,
                                          const double* const* tracers, const obm_pisces_fields* aux, double* const* G,
                                          int accumulate, void* stream) {
    OBM_REQUIRE(row_latitude_daylengths, OBM_ENULL, "obm_pisces_tendencies_rows: the per-row table is NULL");
    static thread_local PiscesArgs tl;
    PiscesArgs& A = tl;
    const int rc = pisces_arguments(A, grid, p, tracers, aux, G, accumulate);
    if (rc) return rc;
    const dim3 gr = cell_grid(A.d, PB);
    if (A.accumulate) pisces_tendency_rows_kernel<true><<<gr, PB, 0, (cudaStream_t)stream>>>(A, row_latitude_daylengths);
    else pisces_tendency_rows_kernel<false><<<gr, PB, 0, (cudaStream_t)stream>>>(A, row_latitude_daylengths);
    return launch_status("pisces_tendency_rows_kernel");
}
