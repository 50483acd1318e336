"""CPU: the RK3 / Euler tracer-update oracle (src/BoxModel/timesteppers.jl:20-28,66-93) against the formula
evaluated in numpy with the same operation order, first-stage form, caching and sub-range behaviour."""
import numpy as np


def grid_and_fields(oracle, nf=3, seed=0):
    rng = np.random.default_rng(seed)
    zf = np.linspace(-8.0, 5.0, 14)
    g = oracle.Grid(5, 4, 7, 2, 1, 3, 0.5 * (zf[:-1] + zf[1:]), zf)
    mk = lambda: [rng.normal(size=g.parent_shape) for _ in range(nf)]  # noqa: E731
    return g, mk(), mk(), mk()


def test_substep_formula_and_cache(oracle):
    g, U, Gn, Gm = grid_and_fields(oracle)
    U0, Gm0 = [u.copy() for u in U], [m.copy() for m in Gm]
    dt, gamma, zeta = 1200.0, 5 / 12, -17 / 60
    oracle.rk3_substep(g, U, Gn, Gm, dt, gamma, zeta, cache_previous=True)
    for f in range(3):
        want = g.interior(U0[f]) + dt * (gamma * g.interior(Gn[f]) + zeta * g.interior(Gm0[f]))
        assert np.array_equal(g.interior(U[f]), want)
        assert np.array_equal(g.interior(Gm[f]), g.interior(Gn[f]))  # G⁻ ← Gⁿ
        halo = U[f].copy()
        g.interior(halo)[...] = g.interior(U0[f])
        assert np.array_equal(halo, U0[f])  # halos untouched


def test_first_stage_and_euler(oracle):
    g, U, Gn, Gm = grid_and_fields(oracle, seed=1)
    U0, Gm0 = [u.copy() for u in U], [m.copy() for m in Gm]
    dt, gamma = 37.5, 8 / 15
    oracle.rk3_substep(g, U, Gn, Gm, dt, gamma, None, cache_previous=False)
    for f in range(3):
        assert np.array_equal(g.interior(U[f]), g.interior(U0[f]) + dt * gamma * g.interior(Gn[f]))  # (Δt·γ¹)·G¹
        assert np.array_equal(Gm[f], Gm0[f])
    # forward Euler is γ = 1 without a ζ term
    V = [u.copy() for u in U0]
    oracle.rk3_substep(g, V, Gn, Gm, dt, 1.0, None, cache_previous=False)
    assert np.array_equal(g.interior(V[0]), g.interior(U0[0]) + dt * g.interior(Gn[0]))


def test_three_stages_integrate_exponential_decay(oracle):
    # dU/dt = −λU with Oceananigans' RK3 coefficients: third-order accurate
    zf = np.array([-1.0, 0.0])
    g = oracle.Grid(1, 1, 1, 0, 0, 0, np.array([-0.5]), zf)
    lam, T = 0.7, 1.0
    errs = []
    for n in (20, 40):
        U, Gn, Gm = [np.ones(g.parent_shape)], [np.zeros(g.parent_shape)], [np.zeros(g.parent_shape)]
        dt = T / n
        Gn[0][...] = -lam * U[0]
        for _ in range(n):
            for gamma, zeta in ((8 / 15, None), (5 / 12, -17 / 60), (3 / 4, -5 / 12)):
                oracle.rk3_substep(g, U, Gn, Gm, dt, gamma, zeta, cache_previous=True)
                Gn[0][...] = -lam * U[0]
        errs.append(abs(U[0].item() - np.exp(-lam * T)))
    assert errs[0] < 1e-6 and errs[0] / errs[1] > 6.5  # ≈ 2³
