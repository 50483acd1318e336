// fused_host.cpp — MEASUREMENT AND TEST INFRASTRUCTURE, NOT PRODUCT.  The GPU kernels' FUSED algorithm compiled for the host.
//
// BASELINE.md §3 promises two CPU columns next to every GPU number: the reference's algorithm and launch structure (the
// oracle: one pass per tracer, the reference's damped Newton) and this one — the algorithm the kernels run (one pass over
// the cells, every shared sub-model evaluated once, all bands in one scan, fixed-iteration ln[H⁺] Newton from the analytic
// starting point) — so that the algorithmic gain and the hardware gain can be told apart.  The arithmetic is NOT restated:
// the very headers the CUDA kernels are built from are included here (csrc/pisces_cell.cuh, csrc/carbon_chemistry.cuh; the
// few device intrinsics they use have plain-C++ shims in csrc/obm_common.cuh, the reciprocal is IEEE 1/x on the host).
// tests/test_fused_host.py also uses it to check the kernel source's operation order against the oracle WITHOUT a GPU.
//
// Only bench.py's `cpu_baseline_fused` leg and tests/ load this library; the product package cannot reach it.
// Build: bench_ref/Makefile (g++ -O2 -ffp-contract=fast -fopenmp — the kernels contract a·b + c into FMAs, so does this).
#define _GNU_SOURCE 1
#include <math.h>
#include <stdarg.h>

static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __all_sync(unsigned, int p) { return p; }   // a "warp" of one lane
static inline int __any_sync(unsigned, int p) { return p; }

#include "../oceanbiome.jl_b200/csrc/pisces_cell.cuh"
#include "../oceanbiome.jl_b200/csrc/carbon_chemistry.cuh"

namespace obm {
void set_error(const char*, ...) {}
int launch_status(const char*) { return 0; }
}  // namespace obm
using namespace obm;

namespace {
struct StoreSink {
    double* const* G;
    long long idx;
    int accumulate;
    inline void put(int n, double t) {
        if (!G[n]) return;
        if (accumulate) G[n][idx] += t; else G[n][idx] = t;
    }
};

inline long long cidx(const GridDims& d, int i, int j, int k) { return (long long)(i + d.Hx) + d.sy * (j + d.Hy) + d.sz * (k + d.Hz); }
inline long long pidx(const GridDims& d, int i, int j) { return (long long)(i + d.Hx) + d.sy * (j + d.Hy); }

// the kernel's load_inputs (pisces_tendencies.cu), with plain loads
inline Inputs load_inputs(const PiscesArgs& a, long long idx, long long pl, int k) {
    Inputs in;
    auto ld = [](const double* p) { return *p; };
    in.P = ld(a.c[T_P] + idx); in.PChl = ld(a.c[T_PChl] + idx); in.PFe = ld(a.c[T_PFe] + idx);
    in.D = ld(a.c[T_D] + idx); in.DChl = ld(a.c[T_DChl] + idx); in.DFe = ld(a.c[T_DFe] + idx); in.DSi = ld(a.c[T_DSi] + idx);
    in.Z = ld(a.c[T_Z] + idx); in.M = ld(a.c[T_M] + idx); in.DOC = ld(a.c[T_DOC] + idx);
    in.POC = ld(a.c[T_POC] + idx); in.GOC = ld(a.c[T_GOC] + idx); in.SFe = ld(a.c[T_SFe] + idx); in.BFe = ld(a.c[T_BFe] + idx);
    in.PSi = ld(a.c[T_PSi] + idx); in.CaCO3 = ld(a.c[T_CaCO3] + idx);
    in.c.NO3 = ld(a.c[T_NO3] + idx); in.c.NH4 = ld(a.c[T_NH4] + idx); in.c.PO4 = ld(a.c[T_PO4] + idx); in.c.Fe = ld(a.c[T_Fe] + idx);
    in.c.Si = ld(a.c[T_Si] + idx); in.c.O2 = ld(a.c[T_O2] + idx); in.c.T = ld(a.c[T_T] + idx);
    in.c.PAR1 = ld(a.f.PAR1 + idx); in.c.PAR2 = ld(a.f.PAR2 + idx); in.c.PAR3 = ld(a.f.PAR3 + idx);
    in.PARt = ld(a.f.PAR + idx); in.Omega = ld(a.f.Omega + idx);
    in.wPOC = (ld(a.f.wPOC + idx) + ld(a.f.wPOC + idx + a.d.sz)) / 2;
    in.wGOC = (ld(a.f.wGOC + idx) + ld(a.f.wGOC + idx + a.d.sz)) / 2;
    in.c.zmxl = a.f.mixed_layer_depth_xy[pl];
    in.c.zeu = a.f.euphotic_depth_xy[pl];
    in.c.kappa = a.f.mean_mixed_layer_vertical_diffusivity_xy[pl];
    in.mlPAR = a.f.mean_mixed_layer_light_xy[pl];
    in.c.z = a.d.zc[k];
    return in;
}
}  // namespace

// The 24 PISCES tendencies of every cell in one pass: cell_tendencies<EXACT> of csrc/pisces_cell.cuh — the kernel's own code.
// exact = 1: the exact pass's arithmetic policy (IEEE division, NaN-propagating min / max), 0: the fast pass's.
extern "C" int fused_pisces_tendencies(const obm_grid* grid, const obm_pisces_params* p, const double* const* tracers,
                                       const obm_pisces_fields* aux, double* const* G, int accumulate, int exact) {
    static thread_local PiscesArgs A;
    memset(&A, 0, sizeof(A));
    int rc = make_dims(grid, &A.d, true);
    if (rc) return rc;
    A.f = *aux;
    for (int n = 0; n < OBM_PISCES_NTRACERS; n++) {
        A.c[n] = tracers[n];
        A.g[n] = (n == T_T || n == T_S) ? nullptr : G[n];
    }
    pisces_prepare(A, p);
    const PiscesArgs& a = A;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < a.d.Nz; k++)
        for (int j = a.d.j0; j < a.d.j1; j++)
            for (int i = a.d.i0; i < a.d.i1; i++) {
                const long long idx = cidx(a.d, i, j, k);
                const Inputs in = load_inputs(a, idx, pidx(a.d, i, j), k);
                StoreSink sink{a.g, idx, accumulate};
                if (exact) cell_tendencies<true>(a, in, sink);
                else cell_tendencies<false>(a, in, sink);
            }
    return 0;
}

// PISCES stage prologue, fused as scale_negative_calcite_kernel does it: all conserved groups on the cell's values, then Ω
// of the same cell from the rescaled DIC, Alk, Si with cc::solve (analytic start, ≤ 12 Newton steps in ln[H⁺]).
extern "C" int fused_scale_negative_tracers_calcite_saturation(const obm_grid* grid, int ntracers, double* const* tracers, int ngroups,
                                                               const obm_scale_group* groups, double fill, const double* T,
                                                               const double* S, const double* DIC, const double* Alk,
                                                               const double* Si, double* Omega, int level_tables) {
    GridDims d;
    int rc = make_dims(grid, &d, true);
    if (rc) return rc;
    int iDIC = -1, iAlk = -1, iSi = -1;
    for (int t = 0; t < ntracers; t++) {
        if (tracers[t] == DIC) iDIC = t;
        if (tracers[t] == Alk) iAlk = t;
        if (tracers[t] == Si) iSi = t;
    }
    // what the first threads of a block build in shared memory: the level's TEOS-10 and pressure-correction tables
    cc::LevelTables* tables = new cc::LevelTables[d.Nz];
    for (int k = 0; k < d.Nz; k++)
        for (int n = 0; n < 64; n++) cc::fill_level_entry(tables[k], fabs(d.zc[k]) * 9.80665 * 1026.0 / 100000.0, n);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < d.Nz; k++)
        for (int j = d.j0; j < d.j1; j++)
            for (int i = d.i0; i < d.i1; i++) {
                const long long idx = cidx(d, i, j, k);
                double v[OBM_MAX_SCALE_TRACERS];
                bool touched = false;
                for (int t = 0; t < ntracers; t++) {
                    v[t] = tracers[t][idx];
                    touched |= (unsigned)__double2hiint(v[t]) >= 0x7ff00000u;
                }
                if (d.bottom != nullptr && (long long)k + 1 < d.bottom[pidx(d, i, j)]) touched = false;  // immersed cell
                if (touched) {  // negative_tracers.cu scale_cell
                    unsigned dirty = 0;
                    for (int q = 0; q < ngroups; q++) {
                        const obm_scale_group& g = groups[q];
                        double t = 0.0, p = 0.0;
                        unsigned members = 0;
                        bool bad = false;
                        for (int m = 0; m < g.n; m++) {
                            const double x = v[g.index[m]], s = x * g.scalefactor[m];
                            t += s;
                            if (x > 0) p += s;
                            bad |= (unsigned)__double2hiint(x) >= 0x7ff00000u;
                            members |= 1u << g.index[m];
                        }
                        if (!bad) continue;
                        t = t < 0 ? fill : t;
                        const double ratio = t / p;
                        for (int m = 0; m < g.n; m++) {
                            const double x = v[g.index[m]];
                            v[g.index[m]] = (!isfinite(x) | (x > 0)) ? x * ratio : 0.0;
                        }
                        dirty |= members;
                    }
                    for (int t = 0; t < ntracers; t++)
                        if ((dirty >> t) & 1u) tracers[t][idx] = v[t];
                }
                const double P = fabs(d.zc[k]) * 9.80665 * 1026.0 / 100000.0;
                Omega[idx] = cc::solve<true>(OBM_CC_OMEGA_CALCITE, T[idx], S[idx], iDIC >= 0 ? v[iDIC] : DIC[idx], iAlk >= 0 ? v[iAlk] : Alk[idx],
                                             P, true, iSi >= 0 ? v[iSi] : Si[idx], false, 0.0, false, 0.0, 1e-8, 12, nullptr,
                                             level_tables ? &tables[k] : nullptr);
            }
    delete[] tables;
    return 0;
}

// All PAR bands of a column in ONE top-down pass (multi_band.jl:147-163 per band), the total, and PISCES' two column
// diagnostics of the total in the same pass (compute_euphotic_depth.jl:3-29, mean_mixed_layer_properties.jl:25-49) — what
// par_multiband_kernel<NB, DIAG> fuses.
extern "C" int fused_par_multiband_column_state(const obm_grid* grid, const obm_multiband_params* p, const double* chl_a,
                                                const double* chl_b, double chl_scale, double surface, double* const* bands,
                                                double* total, const double* zmxl_xy, double cutoff, double* zeu_xy,
                                                double* mean_xy) {
    GridDims d;
    int rc = make_dims(grid, &d, true);
    if (rc) return rc;
    const int Nz = d.Nz, nb = p->nbands;
#pragma omp parallel for schedule(static)
    for (int j = d.j0; j < d.j1; j++)
        for (int i = d.i0; i < d.i1; i++) {
            double par[OBM_MAX_BANDS];
            const long long pl = pidx(d, i, j);
            const double zmxl = zmxl_xy[pl];
            double acc = 0, depth = 0, zeu = -INFINITY, above = 0, surface_PAR = 0;
            for (int k = Nz - 1; k >= 0; k--) {
                const long long idx = cidx(d, i, j, k);
                const double chl = (chl_a[idx] + (chl_b ? chl_b[idx] : 0.0)) * chl_scale;
                double tot = 0;
                for (int b = 0; b < nb; b++) {
                    const double att = p->water_attenuation_coefficient[b] + p->chlorophyll_attenuation_coefficient[b] * pow(chl, p->chlorophyll_exponent[b]);
                    par[b] = k == Nz - 1 ? surface * p->surface_PAR_division[b] * exp(d.zc[k] * att)
                                         : par[b] * exp((d.zc[k] - d.zc[k + 1]) * att);
                    bands[b][idx] = par[b];
                    tot = b == 0 ? par[b] : tot + par[b];
                }
                total[idx] = tot;
                if (k == Nz - 1) surface_PAR = (tot + total[cidx(d, i, j, Nz)]) / 2;  // the halo value, as found
                else if (tot <= surface_PAR * cutoff && isinf(zeu))
                    zeu = d.zc[k] + (log(surface_PAR * cutoff) - log(tot)) * (d.zc[k] - d.zc[k + 1]) / (log(tot) - log(above));
                above = tot;
                const double zk = d.zf[k], zk1 = d.zf[k + 1];
                const double dz = zk >= zmxl ? zk1 - zk : (zk1 > zmxl ? zk1 - zmxl : 0);
                acc += tot * dz;
                depth += dz;
            }
            zeu_xy[pl] = isfinite(zeu) ? zeu : d.zc[-1];
            mean_xy[pl] = acc / depth;
        }
    return 0;
}

// The flat carbonate sweep with the kernels' solve (csrc/carbon_chemistry.cu).
extern "C" int fused_carbon_chemistry(long long n, const double* T, const double* S, const double* DIC, const double* Alk,
                                      int output_kind, int iterations, double* out) {
#pragma omp parallel for schedule(static)
    for (long long c = 0; c < n; c++)
        out[c] = cc::solve<false>(output_kind, T[c], S[c], DIC[c], Alk[c], 0.0, false, 0.0, false, 0.0, false, 0.0, 1e-8, iterations, nullptr);
    return 0;
}

// exp_table of csrc/obm_common.cuh over an array (tests/test_fused_host.py: accuracy against libm)
extern "C" int fused_exp_table(long long n, const double* x, double* out, int clamped) {
    for (long long q = 0; q < n; q++) out[q] = clamped ? exp_table_clamped(x[q]) : exp_table(x[q]);
    return 0;
}
