// obm_common.cuh — shared device/host helpers of libobm_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/obm_b200.h"

#ifndef __CUDACC__
// Host build of the arithmetic headers (g++, bench_ref/fused_host.cpp): the few device intrinsics they use, as plain C++.
#include <string.h>
static inline double __longlong_as_double(long long x) { double d; memcpy(&d, &x, 8); return d; }
static inline long long __double_as_longlong(double d) { long long x; memcpy(&x, &d, 8); return x; }
static inline int __double2hiint(double d) { return (int)(__double_as_longlong(d) >> 32); }
static inline int __double2loint(double d) { return (int)(__double_as_longlong(d) & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) {
    return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo));
}
#endif

namespace obm {

// ---- thread-local error string (obm_last_error) -------------------------------------------
void set_error(const char* fmt, ...);
int launch_status(const char* what);  // cudaGetLastError → 0 or positive cudaError_t (+ message)

#define OBM_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            obm::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

// ---- Julia semantics ---------------------------------------------------------
// Julia max/min propagate NaN; CUDA fmax/fmin return the non-NaN operand.
__device__ __forceinline__ double jl_max(double a, double b) {
    double m = fmax(a, b);
    return (a != a || b != b) ? __longlong_as_double(0x7ff8000000000000LL) : m;
}
__device__ __forceinline__ double jl_min(double a, double b) {
    double m = fmin(a, b);
    return (a != a || b != b) ? __longlong_as_double(0x7ff8000000000000LL) : m;
}
// eps(0.0): smallest subnormal. FP64 subnormals are honoured on sm_100a (no FTZ for doubles).
__device__ __forceinline__ double eps0() { return __longlong_as_double(1LL); }
// eps(x) = ulp(x) (NaN for NaN/Inf)
__device__ __forceinline__ double jl_eps(double x) {
    double ax = fabs(x);
    return __longlong_as_double(__double_as_longlong(ax) + 1) - ax;  // Inf: NaN-pattern − Inf = NaN
}

// Lean reciprocal: MUFU.RCP64H seed r₀ (relative error e ≈ 2⁻²³) + ONE cubic step
// r = r₀(1 + e + e²), e = 1 − b·r₀ (error e³ ≈ 2⁻⁶⁹, i.e. ≤ 1 ulp after rounding): 3 dependent DFMA, no branches,
// no slow-path call.  Valid for normal, finite b (0 → Inf → NaN, NaN → NaN: garbage in, NaN out);
// callers that need the IEEE special cases use `/`.
__device__ __forceinline__ double rcp_fast(double b) {
#ifdef __CUDACC__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    const double e = fma(-b, r, 1.0);
    return fma(fma(e, e, e), r, r);
#else
    return 1.0 / b;  // host build: IEEE reciprocal (≤ 0.5 ulp; the device sequence is ≤ 1 ulp)
#endif
}

// Division for the few quotients that feed a cancellation (PISCES free iron: Fe − Fe′ with Fe′ → Fe): the lean reciprocal
// plus one Markstein correction — q = a·r, q′ = q + (a − q·b)·r — which lands on the correctly rounded quotient (the
// residual a − q·b is exact in an FMA), where a·r alone can be 1 ulp off and the subtraction would turn that ulp into
// the whole result.  3 more FP64 instructions than `a * rcp_fast(b)`, still branch-free.
__device__ __forceinline__ double div_cr(double a, double b) {
#ifdef __CUDACC__
    const double r = rcp_fast(b);
    const double q = a * r;
    return fma(fma(-q, b, a), r, q);
#else
    return a / b;
#endif
}
// a·b and a + b rounded separately (never contracted into an FMA): for expressions whose result is a difference of nearly
// equal numbers, where the reference's (Julia: no contraction) rounding sequence decides which noise comes out.
__device__ __forceinline__ double mul_rn(double a, double b) {
#ifdef __CUDACC__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
__device__ __forceinline__ double add_rn(double a, double b) {
#ifdef __CUDACC__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

// KD(x): the double literal x as a constant-bank operand.  ptxas materialises a 64-bit FP immediate with TWO moves
// (UMOV lo / UMOV hi) in front of every use — in the carbonate solve (≈ 150 literals: TEOS-10, the equilibrium
// constants, their pressure corrections) that was more issue slots than the arithmetic itself (805 moves vs 693 FP64
// instructions, static).  A variable template keyed by the bit pattern puts each distinct literal in bank 3 once; it is
// then fetched by one LDC(U).64 — or two neighbours by one LDCU.128 — and the value is bit-identical.
#ifdef __CUDACC__
template <unsigned long long BITS>
static __constant__ double CONST_BANK_DOUBLE = __builtin_bit_cast(double, BITS);
#define KD(x) (::obm::CONST_BANK_DOUBLE<__builtin_bit_cast(unsigned long long, (double)(x))>)
#else
#define KD(x) ((double)(x))
#endif

// Lean exp: k = round(x·log₂e) by the 1.5·2⁵² shift, r = x − k·ln2 (two-term Cody–Waite), e^r = (1 + r) + r²·Q(r) with Q
// the degree-11 Taylor tail Σ r^m/(m+2)! (|r| ≤ ln2/2 ⇒ truncation < 2⁻⁵⁷) as even / odd Horner chains in r² — ONE
// constant per FMA, fetched from the constant bank — then the exponent of the result is advanced by k.  ≈ 26
// instructions against ≈ 45 for the library exp (whose 64-bit immediates cost two moves each); ≤ 2 ulp for |x| < 700;
// anything else (overflow, underflow to subnormals or 0, NaN) takes the library call, out of line behind a rarely
// taken branch.  Used by the PAR scans (issue-bound); measured and NOT adopted in the PISCES tendency kernel (± 0) and
// the carbonate solve (+6 %: 4 more FP64 instructions per call) — an Estrin variant was slower everywhere (two
// constants per pair FMA) and is gone.
#ifndef __CUDACC__
#undef __constant__
#define __constant__
#undef __noinline__
#define __noinline__
#endif
static __constant__ double EXP_C[12] = {  // 1/n!, n = 2 … 13
    0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800,
    1.0 / 479001600, 1.0 / 6227020800.0};
static __device__ __noinline__ double exp_library(double x) { return exp(x); }  // out of line: keeps the hot code short
__device__ __forceinline__ double exp_lean(double x) {
    if (!(fabs(x) < 700.0)) return exp_library(x);
    const double SHIFT = KD(6755399441055744.0);
    const double t = fma(x, KD(1.4426950408889634), SHIFT);
    const int k = __double2loint(t);
    const double kf = t - SHIFT;
    double r = fma(kf, KD(-6.93147180559945286e-01), x);
    r = fma(kf, KD(-2.31904681384629956e-17), r);
    const double r2 = r * r;
    const double a0 = 1.0 + r;
    double e = fma(EXP_C[10], r2, EXP_C[8]);
    double o = fma(EXP_C[11], r2, EXP_C[9]);
    e = fma(e, r2, EXP_C[6]);
    o = fma(o, r2, EXP_C[7]);
    e = fma(e, r2, EXP_C[4]);
    o = fma(o, r2, EXP_C[5]);
    e = fma(e, r2, EXP_C[2]);
    o = fma(o, r2, EXP_C[3]);
    e = fma(e, r2, EXP_C[0]);
    o = fma(o, r2, EXP_C[1]);
    const double p = fma(fma(o, r, e), r2, a0);
    return __hiloint2double(__double2hiint(p) + (int)((unsigned)k << 20), __double2loint(p));
}

// e^x for |x| < 700 (no guard): k = round(x·log₂e) by the 1.5·2⁵² shift, r = x − k·ln2 (two-term Cody–Waite),
// e^r by the degree-13 Taylor polynomial (|r| ≤ ln2/2 ⇒ truncation r¹⁴/14! < 2⁻⁵⁷), exponent advanced by k.  ≤ 2 ulp.
// NaN → NaN (the shifted sum of a NaN is the canonical NaN, low word 0 ⇒ k = 0).
__device__ __forceinline__ double exp_unguarded(double x) {
    const double SHIFT = KD(6755399441055744.0);
    const double t = fma(x, KD(1.4426950408889634), SHIFT);
    const int k = __double2loint(t);
    const double kf = t - SHIFT;
    double r = fma(kf, KD(-6.93147180559945286e-01), x);
    r = fma(kf, KD(-2.31904681384629956e-17), r);
    double p = fma(KD(1.0 / 6227020800.0), r, KD(1.0 / 479001600));
    p = fma(p, r, KD(1.0 / 39916800));
    p = fma(p, r, KD(1.0 / 3628800));
    p = fma(p, r, KD(1.0 / 362880));
    p = fma(p, r, KD(1.0 / 40320));
    p = fma(p, r, KD(1.0 / 5040));
    p = fma(p, r, KD(1.0 / 720));
    p = fma(p, r, KD(1.0 / 120));
    p = fma(p, r, KD(1.0 / 24));
    p = fma(p, r, KD(1.0 / 6));
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (int)((unsigned)k << 20), __double2loint(p));
}
// e^x from a 64-entry table: k = round(x·64/ln2), r = x − k·ln2/64 (two-term Cody–Waite, |r| ≤ ln2/128), e^x =
// 2^(k>>6) · T[k & 63] · e^r with T[j] = 2^(j/64) correctly rounded and e^r − 1 = r(1 + r/2 + r²/6 + r³/24 + r⁴/120)
// (truncation r⁶/720 < 3.5e-17).  10 FP64 instructions and one 8-byte read of a 512-byte table that lives in L1 (read-only
// path) against 17 FP64 instructions + 13 constant-bank operands for the degree-13 polynomial of exp_unguarded; ≤ 1 ulp.
// Valid for −707 ≤ x ≤ 709 (normal results; callers clamp or test the range); NaN → NaN.
#ifdef __CUDACC__
static __device__ const double EXP_T64[64] = {
#else
static const double EXP_T64[64] = {
#endif
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};
__device__ __forceinline__ double exp_table(double x) {
    const double SHIFT = KD(6755399441055744.0);
    const double t = fma(x, KD(92.33248261689366), SHIFT);
    const int k = __double2loint(t);
    const double kf = t - SHIFT;
    double r = fma(kf, KD(-0.010830424667801708), x);
    r = fma(kf, KD(-2.8447437476627285e-11), r);
    double q = fma(KD(1.0 / 120), r, KD(1.0 / 24));
    q = fma(q, r, KD(1.0 / 6));
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    const double p = q * r;  // e^r − 1
#ifdef __CUDACC__
    const double T = __ldg(&EXP_T64[k & 63]);
#else
    const double T = EXP_T64[k & 63];
#endif
    const double v = fma(T, p, T);
    return __hiloint2double(__double2hiint(v) + (int)((unsigned)(k >> 6) << 20), __double2loint(v));
}
// the same behind a clamp of the argument to the valid range: e^x for every finite x up to the clamp's effect (x < −707:
// 8.7e-308 instead of the subnormals / 0 below it; x > 709: 8.2e307 instead of +Inf) — for callers whose small results
// vanish in a sum anyway
__device__ __forceinline__ double exp_table_clamped(double x) {
    return exp_table(x < -707.0 ? -707.0 : (x > 709.0 ? 709.0 : x));
}
// |x| < 700 and not NaN ⇔ the high word without its sign is below that of 700.0 (0x4085e000): one integer compare
__device__ __forceinline__ bool exp_in_range(double x) { return ((unsigned)__double2hiint(x) & 0x7fffffffu) < 0x4085e000u; }
__device__ __forceinline__ double exp_horner(double x) {
    if (!exp_in_range(x)) return exp(x);  // overflow, underflow to subnormals / 0, ±Inf, NaN: the library's answer
    return exp_unguarded(x);
}
// ln x for positive normal x: x = 2^e·m, m ∈ [√½, √2), s = (m − 1)/(m + 1), ln m = 2s·Σ s^(2k)/(2k + 1) to k = 9
// (s² ≤ 0.0295 ⇒ truncation < 2⁻⁵⁵), ln x = e·ln2 + ln m with ln2 split in two.  ≤ 2 ulp; ≈ 30 instructions against
// ≈ 60 for the library log (two of them per solve: ln T and ln(1 − 0.001005 S)).  Anything else (≤ 0, subnormal, ±Inf,
// NaN) takes the library log, inline in the cold branch.
// positive and normal ⇔ one unsigned compare of the high word
__device__ __forceinline__ bool log_in_range(double x) { return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fe00000u; }
__device__ __forceinline__ double log_unguarded(double x) {  // valid for positive normal x only (see log_in_range)
    const int hi = __double2hiint(x);
    int e = (hi >> 20) - 1023;
    int mh = (hi & 0x000fffff) | 0x3ff00000;      // m ∈ [1, 2)
    const bool up = mh >= 0x3ff6a09f;              // m ≥ √2 (to 2⁻²⁰): halve it
    mh = up ? mh - 0x00100000 : mh;
    e = up ? e + 1 : e;
    const double m = __hiloint2double(mh, __double2loint(x));
    const double f = m - 1.0;
    const double s = f * rcp_fast(2.0 + f);
    const double z = s * s;
    double q = fma(KD(2.0 / 19), z, KD(2.0 / 17));
    q = fma(q, z, KD(2.0 / 15));
    q = fma(q, z, KD(2.0 / 13));
    q = fma(q, z, KD(2.0 / 11));
    q = fma(q, z, KD(2.0 / 9));
    q = fma(q, z, KD(2.0 / 7));
    q = fma(q, z, KD(2.0 / 5));
    q = fma(q, z, KD(2.0 / 3));
    const double lm = fma(s * z, q, 2.0 * s);      // ln m
    const double ef = (double)e;
    return fma(ef, KD(6.93147180369123816490e-01), fma(ef, KD(1.90821492927058770002e-10), lm));
}
__device__ __forceinline__ double log_lean(double x) {
    if (!log_in_range(x)) return log(x);
    return log_unguarded(x);
}
// ---- grid indexing -----------------------------------------------------------------------------
struct GridDims {
    int Nx, Ny, Nz, Hx, Hy, Hz;
    int i0, i1, j0, j1;
    long long sy, sz;  // element strides of the parent array (sx = 1)
    const double* zc;  // shifted so that zc[k] is interior level k (k = -Hz … Nz-1+Hz valid)
    const double* zf;
    const long long* bottom;  // nullable: 1-based index of the bottom-most active cell per column (x–y parent plane)
};

inline int make_dims(const obm_grid* g, GridDims* d, bool need_z) {
    OBM_REQUIRE(g != nullptr, OBM_ENULL, "grid is NULL");
    OBM_REQUIRE(g->Nx > 0 && g->Ny > 0 && g->Nz > 0 && g->Hx >= 0 && g->Hy >= 0 && g->Hz >= 0, OBM_ESIZE,
                "bad grid size N=(%d,%d,%d) H=(%d,%d,%d)", g->Nx, g->Ny, g->Nz, g->Hx, g->Hy, g->Hz);
    d->Nx = g->Nx; d->Ny = g->Ny; d->Nz = g->Nz;
    d->Hx = g->Hx; d->Hy = g->Hy; d->Hz = g->Hz;
    d->i0 = g->i0; d->j0 = g->j0;
    d->i1 = g->i1 > 0 ? g->i1 : g->Nx;
    d->j1 = g->j1 > 0 ? g->j1 : g->Ny;
    OBM_REQUIRE(d->i0 >= 0 && d->i1 <= g->Nx && d->i0 < d->i1 && d->j0 >= 0 && d->j1 <= g->Ny && d->j0 < d->j1,
                OBM_ESIZE, "bad sub-range i=[%d,%d) j=[%d,%d)", d->i0, d->i1, d->j0, d->j1);
    d->bottom = (const long long*)g->bottom_indices_xy;
    d->sy = (long long)g->Nx + 2 * g->Hx;
    d->sz = d->sy * ((long long)g->Ny + 2 * g->Hy);
    if (need_z) {
        OBM_REQUIRE(g->zc != nullptr && g->zf != nullptr, OBM_ENULL, "grid.zc / grid.zf is NULL");
        d->zc = g->zc + g->Hz;
        d->zf = g->zf + g->Hz;
    } else {
        d->zc = g->zc ? g->zc + g->Hz : nullptr;
        d->zf = g->zf ? g->zf + g->Hz : nullptr;
    }
    return 0;
}

__device__ __forceinline__ long long cell_index(const GridDims& d, int i, int j, int k) {
    return (long long)(i + d.Hx) + d.sy * (j + d.Hy) + d.sz * (k + d.Hz);
}
__device__ __forceinline__ long long plane_index(const GridDims& d, int i, int j) {
    return (long long)(i + d.Hx) + d.sy * (j + d.Hy);
}

#ifdef __CUDACC__
// One thread per interior cell of the sub-range, x fastest (coalesced along the contiguous axis).
// 3-D launch without any integer division: blockIdx.x ↔ x chunk, blockIdx.y ↔ j, blockIdx.z ↔ k
// (when Ny exceeds the 65535 limit of gridDim.y, j is folded into blockIdx.x).
__device__ __forceinline__ bool thread_cell(const GridDims& d, int& i, int& j, int& k) {
    const int nx = d.i1 - d.i0;
    const unsigned chunks = (nx + blockDim.x - 1) / blockDim.x;
    unsigned bx = blockIdx.x, by = blockIdx.y;
    if (gridDim.y == 1 && (d.j1 - d.j0) > 1) {  // folded layout
        by = bx / chunks;
        bx = bx - by * chunks;
    }
    const int ii = (int)(bx * blockDim.x + threadIdx.x);
    if (ii >= nx) return false;
    i = d.i0 + ii;
    j = d.j0 + (int)by;
    k = (int)blockIdx.z;
    return true;
}
inline dim3 cell_grid(const GridDims& d, int block) {
    const unsigned chunks = (unsigned)((d.i1 - d.i0 + block - 1) / block);
    const unsigned ny = (unsigned)(d.j1 - d.j0);
    if (ny <= 65535u) return dim3(chunks, ny, (unsigned)d.Nz);
    return dim3(chunks * ny, 1, (unsigned)d.Nz);
}
#endif  // __CUDACC__
// `immersed_cell(i, j, k, grid)` of a grid-fitted bottom: below the bottom-most active cell of the column
__device__ __forceinline__ bool immersed_cell(const GridDims& d, int i, int j, int k) {
    return d.bottom != nullptr && (long long)k + 1 < d.bottom[plane_index(d, i, j)];
}
inline long long cell_count(const GridDims& d) { return (long long)(d.i1 - d.i0) * (d.j1 - d.j0) * d.Nz; }
inline long long column_count(const GridDims& d) { return (long long)(d.i1 - d.i0) * (d.j1 - d.j0); }

}  // namespace obm
