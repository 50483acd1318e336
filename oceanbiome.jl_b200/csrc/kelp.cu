// kelp.cu — biologically active particles with the sugar-kelp individual model.
//
// Replaces src/Particles/update_tracer_tendencies.jl:1-48 (one launch per coupled tracer: 8, each re-evaluating
// growth, uptake, respiration, erosion AND the Newton solve for the light-inhibition parameter β), tendencies.jl:3-35
// (one launch per particle field: 3) and time_stepping.jl:29-48 (3 Euler launches) by two launches.  One thread per
// particle: the particle state and the tracers of its nearest cell are read once, every shared sub-expression of
// src/Models/Individuals/SugarKelp/equations.jl is evaluated once, and the scatter launch adds all eight
// uptake / release terms to Gⁿ with no-return atomics (RED.ADD.F64 at the L2; several particles may share a cell).
// N is 10¹–10³ particles, so these kernels are launch-latency sized; the point is 2 launches instead of 14.
#include <string.h>

#include "obm_common.cuh"

namespace obm {

constexpr double KDAY = 86400.0;

struct KelpArgs {
    GridDims d;
    obm_sugar_kelp_params p;
    obm_particles q;
    obm_kelp_tracers f;
    double* G[OBM_KELP_NCOUPLED];
    double* dout[3];
    double seasonal;  // seasonal_limitation(kelp, t): depends on the clock only, evaluated on the host
    double dt;
    int iterations;
};

// get_node — tracer_interpolation.jl:5-7 (1-based in, 0-based out)
__device__ __forceinline__ int get_node(int topo, long long i1, int N) {
    if (topo == OBM_TOPO_FLAT) return 0;
    if (topo == OBM_TOPO_BOUNDED) return (int)(i1 < 1 ? 1 : (i1 > N ? N : i1)) - 1;
    return (int)(i1 < 1 ? N : (i1 > N ? 1 : i1)) - 1;
}

// nearest_node (tracer_interpolation.jl:43-62): the fractional index between cell centres, lower node when its
// fractional part is < 1/2, upper node otherwise, then get_node per topology
__device__ __forceinline__ int nearest_regular(double x, double x0, double dx, int topo, int N) {
    if (topo == OBM_TOPO_FLAT) return 0;
    const double fi = (x - x0) / dx;          // 0 at the first centre
    const double fl = floor(fi);
    const long long lo = (long long)fl + 1;   // 1-based lower node
    return get_node(topo, (fi - fl) < 0.5 ? lo : lo + 1, N);
}
__device__ __forceinline__ int nearest_z(const GridDims& d, double z, int topo) {
    if (topo == OBM_TOPO_FLAT) return 0;
    const int N = d.Nz;
    // centres are increasing; find lo = last centre ≤ z (−1 when below the first), fractional part between centres
    int lo = -1, hi = N;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (d.zc[mid] <= z) lo = mid; else hi = mid;
    }
    // halo centres exist on both sides (Hz ≥ 1 is not required: clamp the interval used for the fraction)
    const double zl = d.zc[lo < 0 ? 0 : lo], zu = d.zc[lo + 1 > N - 1 ? N - 1 : lo + 1];
    double frac;
    if (lo < 0) frac = 1.0;            // below the first centre → node 1 after the bounded clamp
    else if (lo >= N - 1) frac = 0.0;  // above the last centre → node N
    else frac = (z - zl) / (zu - zl);
    const long long lo1 = (long long)lo + 1;
    return get_node(topo, frac < 0.5 ? lo1 : lo1 + 1, N);
}

struct KelpState { double A, N, C, u, v, w, T, NO3, NH4, PAR; };
struct KelpRates {
    double mu, nu, J_NO3, J_NH4, P, R, e;  // growth, erosion, uptakes, photosynthesis, respiration, carbon exudate fraction
};

// maximum_photosynthesis(α, β), its β-derivative — equations.jl:119-121
__device__ __forceinline__ double max_photo(double a, double b) {
    return a / (log(1 + a / b)) * (a / (a + b)) * pow(b / (a + b), b / a);
}
__device__ __forceinline__ double d_max_photo(double a, double b) {
    const double L = log(a / b + 1);
    return (a * pow(b / (b + a), b / a) * ((L * (b * b) + a * L * b) * log(b / (b + a)) + a * a)) / (L * L * b * ((b + a) * (b + a)));
}

// Every rate of equations.jl evaluated once, in the reference's operation order.
__device__ __forceinline__ KelpRates kelp_rates(const obm_sugar_kelp_params& k, const KelpState& s, double seasonal, int iterations) {
    KelpRates r;
    // current_factor :185-193
    const double U = sqrt(s.u * s.u + s.v * s.v + s.w * s.w);
    const double fc = k.current_1 * (1 - exp(-U / k.current_3)) + k.current_2;
    // potential_ammonia_uptake :86-93
    const double jNH4 = k.maximum_ammonia_uptake * fc * s.NH4 / (k.ammonia_half_saturation + s.NH4);
    // base_growth_limitation :195-227: temperature · area · season
    const double Tl = k.lower_optimal, Tu = k.upper_optimal;
    const double fT = jl_max(0.0, k.lower_gradient * (s.T - Tl) + 1) * (s.T < Tl ? 1.0 : 0.0)
                      + jl_max(0.0, k.upper_gradient * (s.T - Tu) + 1) * (s.T > Tu ? 1.0 : 0.0)
                      + ((Tl <= s.T && s.T <= Tu) ? 1.0 : 0.0);
    const double A0 = k.growth_rate_adjustment;
    const double fA = k.growth_adjustment_1 * exp(-((s.A / A0) * (s.A / A0))) + k.growth_adjustment_2;
    const double f = fT * fA * seasonal;
    // growth :38-58
    const double kA = k.structural_dry_weight_per_area, Ns = k.structural_nitrogen, Cs = k.structural_carbon;
    const double muNH4 = jNH4 / kA / (s.N + Ns);
    const double muN = 1 - k.minimum_nitrogen_reserve / s.N;
    const double muC = 1 - k.minimum_carbon_reserve / s.C;
    r.mu = f * jl_min(muC, jl_max(muN, muNH4));
    // erosion :178-183
    const double ee = exp(k.erosion_exponent * s.A);
    r.nu = k.base_erosion_rate * ee / (1 + k.base_erosion_rate * (ee - 1));
    // nitrate_uptake :60-71, ammonia_uptake :73-84
    const double Nmax = k.maximum_nitrogen_reserve, Nmin = k.minimum_nitrogen_reserve;
    r.J_NO3 = jl_max(0.0, k.maximum_nitrate_uptake * fc * (Nmax - s.N) / (Nmax - Nmin) * s.NO3 / (k.nitrate_half_saturation + s.NO3));
    r.J_NH4 = jl_min(jNH4, r.mu * kA * (s.N + Ns));
    // photosynthesis :95-117 (Tₚₗ is photosynthesis_ref_temp_1 there, not photosynthesis_low_temp — kept)
    {
        const double PAR = s.PAR * (KDAY / (3.99e-10 * 545e12));
        const double Tk = s.T + 273.15;
        const double Ta = k.photosynthesis_arrhenius_temp, Tal = k.photosynthesis_low_arrhenius_temp, Tah = k.photosynthesis_high_arrhenius_temp;
        const double Tp = k.photosynthesis_ref_temp_1, Tpl = k.photosynthesis_ref_temp_1, Tph = k.photosynthesis_high_temp;
        const double a = k.photosynthetic_efficiency, Is = k.saturation_irradiance;
        const double Pm = k.photosynthesis_at_ref_temp_1 * exp(Ta / Tp - Ta / Tk) / (1 + exp(Tal / Tk - Tal / Tpl) + exp(Tah / Tph - Tah / Tk));
        // solve_for_light_inhibition :110-117 + NewtonRaphsonSolver (Utils/solvers.jl:6-22) from β₀ = 1e-9.  The
        // reference's tolerance (eps(1e-9) ≈ 2e-25) is below the residual's rounding noise, so it iterates 1000
        // times; here the iteration stops when the step no longer changes β (same root to rounding).
        double b = 1e-9;
        const double target = Pm / Is;
        for (int n = 0; n < iterations; n++) {
            const double fx = max_photo(a, b) - target;
            const double step = fx / d_max_photo(a, b);
            const double nb = b - step;
            const bool done = !(fabs(step) > 4e-16 * fabs(nb));
            b = nb;
            if (done) break;
        }
        const double ps = a * Is / log(1 + a / b);
        r.P = ps * (1 - exp(-a * PAR / ps)) * exp(-b * PAR / ps);
    }
    // specific_carbon_exudate :160-165
    r.e = 1 - exp(k.exudation * (k.minimum_carbon_reserve - s.C));
    // respiration :135-158
    {
        const double Tk = s.T + 273.15;
        const double fR = exp(k.respiration_arrhenius_temp / k.respiration_ref_temp_1 - k.respiration_arrhenius_temp / Tk);
        const double Jm = k.maximum_nitrate_uptake + k.maximum_ammonia_uptake;
        const double J = r.J_NO3 + r.J_NH4;
        r.R = fR * (k.base_basal_respiration_rate + k.base_activity_respiration_rate * (r.mu / k.maximum_specific_growth_rate + J / Jm));
    }
    (void)Cs;
    return r;
}

__device__ __forceinline__ bool load_particle(const KelpArgs& a, long long n, KelpState& s, long long& idx, double& volume) {
    const obm_particles& q = a.q;
    const int i = nearest_regular(q.x[n], q.x0, q.dx, q.topology[0], a.d.Nx);
    const int j = nearest_regular(q.y[n], q.y0, q.dy, q.topology[1], a.d.Ny);
    const int k = nearest_z(a.d, q.z[n], q.topology[2]);
    idx = cell_index(a.d, i, j, k);
    volume = q.dx * q.dy * (a.d.zf[k + 1] - a.d.zf[k]);
    s.A = q.A[n]; s.N = q.N[n]; s.C = q.C[n];
    s.u = a.f.u ? a.f.u[idx] : 0.0; s.v = a.f.v ? a.f.v[idx] : 0.0; s.w = a.f.w ? a.f.w[idx] : 0.0;
    s.T = a.f.T[idx]; s.NO3 = a.f.NO3[idx]; s.NH4 = a.f.NH4[idx]; s.PAR = a.f.PAR[idx];
    return true;
}

// coupling.jl:3-57 — all eight terms of a particle from one evaluation, scattered atomically
__global__ void __launch_bounds__(128) kelp_scatter_kernel(const __grid_constant__ KelpArgs a) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.q.n) return;
    KelpState s;
    long long idx;
    double volume;
    load_particle(a, n, s, idx, volume);
    const obm_sugar_kelp_params& k = a.p;
    const KelpRates r = kelp_rates(k, s, a.seasonal, a.iterations);
    const double kA = k.structural_dry_weight_per_area;
    const double perN = KDAY * 14 * 0.001, perC = KDAY * 12 * 0.001;  // g N (C) dm⁻² day⁻¹ → mmol N (C) s⁻¹
    double t[OBM_KELP_NCOUPLED];
    t[0] = -r.J_NO3 * s.A / perN;                                        // NO₃
    t[1] = -r.J_NH4 * s.A / perN;                                        // NH₄
    t[2] = -(r.P - r.R) * s.A / perC;                                    // DIC
    t[3] = -t[2];                                                        // O₂
    t[4] = r.e * r.P * s.A / perC;                                       // DOC
    t[5] = t[4] / k.exudation_redfield_ratio;                            // DON
    t[6] = r.nu * kA * s.A * (s.C + k.structural_carbon) / perC;         // bPOC
    t[7] = r.nu * kA * s.A * (s.N + k.structural_nitrogen) / perN;       // bPON
    const double sf = a.q.scalefactors ? a.q.scalefactors[n] : 1.0;
#pragma unroll
    for (int c = 0; c < OBM_KELP_NCOUPLED; c++)
        if (a.G[c]) atomicAdd(a.G[c] + idx, sf * t[c] / volume);  // result unused ⇒ RED.ADD.F64
}

// equations.jl:1-36 + time_stepping.jl:45-48
__global__ void __launch_bounds__(128) kelp_step_kernel(const __grid_constant__ KelpArgs a) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.q.n) return;
    KelpState s;
    long long idx;
    double volume;
    load_particle(a, n, s, idx, volume);
    const obm_sugar_kelp_params& k = a.p;
    const KelpRates r = kelp_rates(k, s, a.seasonal, a.iterations);
    const double kA = k.structural_dry_weight_per_area;
    const double dA = s.A * (r.mu - r.nu) / KDAY;
    const double eN = r.P * r.e * 14 / 12 / k.exudation_redfield_ratio;  // nitrogen_exudate :167-176
    const double dN = (((r.J_NO3 + r.J_NH4) - eN) / kA - r.mu * (s.N + k.structural_nitrogen)) / KDAY;
    const double dC = ((r.P * (1 - r.e) - r.R) / kA - r.mu * (s.C + k.structural_carbon)) / KDAY;
    if (a.dout[0]) { a.dout[0][n] = dA; a.dout[1][n] = dN; a.dout[2][n] = dC; }
    a.q.A[n] = s.A + dA * a.dt;
    a.q.N[n] = s.N + dN * a.dt;
    a.q.C[n] = s.C + dC * a.dt;
}

// day_length, normed_day_length_change, seasonal_limitation — equations.jl:229-255 (host: a function of the clock only)
static double kelp_day_length(double phi, double n) {
    const double PI = 3.14159265358979323846;
    n -= 171;
    double M = fmod(356.5291 + 0.98560028 * n, 360.0);
    if (M < 0) M += 360.0;  // Julia mod: result has the sign of the divisor
    const double Cc = 1.9148 * sin(M * PI / 180) + 0.02 * sin(2 * M * PI / 180) + 0.0003 * sin(3 * M * PI / 180);
    double lam = fmod(M + Cc + 180 + 102.9372, 360.0);
    if (lam < 0) lam += 360.0;
    const double delta = asin(sin(lam * PI / 180) * sin(23.44 * PI / 180));
    const double omega = (sin(-0.83 * PI / 180) * sin(phi * PI / 180) * sin(delta)) / (cos(phi * PI / 180) * cos(delta));
    return omega / 180;
}
static double kelp_seasonal_limitation(const obm_sugar_kelp_params* k, double t) {
    double m = fmod(t, 364 * KDAY);
    if (m < 0) m += 364 * KDAY;
    const double n = floor(m / KDAY);
    const double phi = k->adapted_latitude;
    const double lam = (kelp_day_length(phi, n) - kelp_day_length(phi, n - 1)) / (kelp_day_length(phi, 76) - kelp_day_length(phi, 75));
    const double sg = lam > 0 ? 1.0 : (lam < 0 ? -1.0 : lam);
    return k->photoperiod_1 * (1 + sg * sqrt(fabs(lam))) + k->photoperiod_2;
}

static int fill(KelpArgs& a, const char* who, const obm_grid* grid, const obm_sugar_kelp_params* p, const obm_particles* q,
                const obm_kelp_tracers* f, double t) {
    OBM_REQUIRE(p && q && f, OBM_ENULL, "%s: params / particles / tracers is NULL", who);
    OBM_REQUIRE(q->n >= 0, OBM_ESIZE, "%s: n = %lld", who, (long long)q->n);
    memset(&a, 0, sizeof(a));
    int rc = make_dims(grid, &a.d, true);
    if (rc) return rc;
    if (q->n == 0) return 0;
    OBM_REQUIRE(q->x && q->y && q->z && q->A && q->N && q->C, OBM_ENULL, "%s: a particle array is NULL", who);
    OBM_REQUIRE(f->T && f->NO3 && f->NH4 && f->PAR, OBM_ENULL, "%s: T / NO3 / NH4 / PAR is NULL", who);
    for (int c = 0; c < 3; c++)
        OBM_REQUIRE(q->topology[c] >= OBM_TOPO_PERIODIC && q->topology[c] <= OBM_TOPO_FLAT, OBM_EENUM, "%s: unknown topology %d", who,
                    q->topology[c]);
    OBM_REQUIRE((q->topology[0] == OBM_TOPO_FLAT || q->dx > 0) && (q->topology[1] == OBM_TOPO_FLAT || q->dy > 0), OBM_ESIZE,
                "%s: dx = %g, dy = %g", who, q->dx, q->dy);
    a.p = *p; a.q = *q; a.f = *f;
    a.seasonal = kelp_seasonal_limitation(p, t);
    a.iterations = p->newton_iterations > 0 ? p->newton_iterations : 100;
    return 0;
}

}  // namespace obm

using namespace obm;

extern "C" int obm_kelp_update_tendencies(const obm_grid* grid, const obm_sugar_kelp_params* p, const obm_particles* particles,
                                          const obm_kelp_tracers* tracers, double* const* G, double t, void* stream) {
    static thread_local KelpArgs a;
    int rc = fill(a, "obm_kelp_update_tendencies", grid, p, particles, tracers, t);
    if (rc || particles->n == 0) return rc;
    OBM_REQUIRE(G != nullptr, OBM_ENULL, "obm_kelp_update_tendencies: G is NULL");
    for (int c = 0; c < OBM_KELP_NCOUPLED; c++) a.G[c] = G[c];
    kelp_scatter_kernel<<<(unsigned)((particles->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("kelp_scatter_kernel");
}

extern "C" int obm_kelp_step(const obm_grid* grid, const obm_sugar_kelp_params* p, const obm_particles* particles,
                             const obm_kelp_tracers* tracers, double t, double dt, double* const* tendencies_out,
                             void* stream) {
    static thread_local KelpArgs a;
    int rc = fill(a, "obm_kelp_step", grid, p, particles, tracers, t);
    if (rc || particles->n == 0) return rc;
    a.dt = dt;
    if (tendencies_out) {
        OBM_REQUIRE(tendencies_out[0] && tendencies_out[1] && tendencies_out[2], OBM_ENULL, "obm_kelp_step: a tendency array is NULL");
        for (int c = 0; c < 3; c++) a.dout[c] = tendencies_out[c];
    }
    kelp_step_kernel<<<(unsigned)((particles->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status("kelp_step_kernel");
}

// seasonal_limitation(kelp, t) as the kernels use it (exposed for the tests / host mirror)
extern "C" double obm_kelp_seasonal_limitation(const obm_sugar_kelp_params* p, double t) {
    return p ? kelp_seasonal_limitation(p, t) : NAN;
}
