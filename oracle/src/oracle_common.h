/*
 * oracle_common.h — shared helpers of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a plain-C, scalar, operation-order-faithful
 * restatement of OceanBioME.jl v0.17.6's arithmetic for the biogeochemical hot path.  It is
 * imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs, never by the product path (oceanbiome.jl_b200/), which fails loudly when the
 * CUDA library is missing.
 *
 * Build: -O2 -ffp-contract=off -fno-fast-math  (Julia does not contract a*b+c, does not
 * reassociate, honours subnormals — SURVEY App. A).
 *
 * The reference itself cannot run here (no Julia in the image, Oceananigans /
 * SeawaterPolynomials not on disk), so the oracle is pinned against the reference's own
 * known answers — see tests/test_oracle_*.py and tests/golden/.
 */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/obm_b200.h"

/* Julia `max`/`min` propagate NaN; C fmax/fmin do not. */
static inline double jl_max(double a, double b) {
#ifndef ORC_SELECT_MINMAX /* defined only for _build/liboracle_select.so: compare + select, what the kernels' fast pass does */
    if (isnan(a) || isnan(b)) return NAN;
#endif
    return a > b ? a : b;
}
static inline double jl_min(double a, double b) {
#ifndef ORC_SELECT_MINMAX
    if (isnan(a) || isnan(b)) return NAN;
#endif
    return a < b ? a : b;
}
/* Julia `eps(x)` = ulp(x): 5e-324 at 0, NaN at NaN/Inf. */
static inline double jl_eps(double x) {
    double ax = fabs(x);
    return nextafter(ax, INFINITY) - ax;
}
#define EPS0 4.9406564584124654e-324 /* eps(0.0) */

static inline double jl_sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }

/* parent-array linear index of interior cell (i,j,k), 0-based */
static inline int64_t cell_index(const obm_grid* g, int i, int j, int k) {
    int64_t sy = (int64_t)g->Nx + 2 * g->Hx;
    int64_t sz = sy * ((int64_t)g->Ny + 2 * g->Hy);
    return (i + g->Hx) + sy * (j + g->Hy) + sz * (k + g->Hz);
}
static inline int64_t plane_index(const obm_grid* g, int i, int j) {
    int64_t sy = (int64_t)g->Nx + 2 * g->Hx;
    return (i + g->Hx) + sy * (j + g->Hy);
}
static inline void grid_range(const obm_grid* g, int* i0, int* i1, int* j0, int* j1) {
    *i0 = g->i0;
    *j0 = g->j0;
    *i1 = g->i1 > 0 ? g->i1 : g->Nx;
    *j1 = g->j1 > 0 ? g->j1 : g->Ny;
}

/* Parity metric scale: S = Σ |additive terms| of the tendency this thread evaluated last.  Every per-tracer
 * callable of oracle_pisces.c / oracle_npd.c records it next to its return statement through ORC_TERMS (the return
 * expression itself is untouched: same operations, same order).  A term that is itself a difference of fluxes
 * contributes its own Σ|terms| (read back from orc_term_scale right after the call). */
extern __thread double orc_term_scale;
static inline double orc_abs_sum(const double* v, int n) {
    double s = 0;
    for (int i = 0; i < n; i++) s += fabs(v[i]);
    return s;
}
#define ORC_TERMS(...) \
    (orc_term_scale = orc_abs_sum((const double[]){__VA_ARGS__}, (int)(sizeof((const double[]){__VA_ARGS__}) / sizeof(double))))

#endif
