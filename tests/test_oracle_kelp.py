"""CPU: the sugar-kelp / particle oracle (oracle_kelp.c).  The reference holds no absolute kelp tendency; what pins the
restatement is (a) the bookkeeping between the particle equations (equations.jl:1-36) and the tracer coupling
(coupling.jl:3-57) that makes test/test_sugar_kelp.jl's nitrogen and carbon conservation hold, (b) the defining equation
of the light-inhibition parameter, (c) the documented parameter values, (d) the nearest-node deposit rule the reference's
particle tests state (test/test_particles.jl:97-114), and — r02 — (e) an independent plain-Python transliteration of
equations.jl / coupling.jl (oracle/pyref_kelp.py, no code shared with oracle_kelp.c) through committed golden vectors
(tests/golden/kelp_rates.json by scripts/make_kelp_golden.py): absolute rates, all eleven of them."""
import math

import numpy as np
import pytest

import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib

day = 86400.0


def params(**kw):
    return ob.SugarKelp(**kw).c_params()


def test_derived_parameter_defaults():
    k = ob.SugarKelp()
    # SugarKelp.jl:88-139
    assert k.photosynthetic_efficiency == pytest.approx(4.15e-5 * 24e6 / 86400)
    assert k.saturation_irradiance == pytest.approx(90 * day / 1e6)
    r = 1 - 0.0126 / 0.0216
    assert k.growth_adjustment_2 == pytest.approx(0.039 / (2 * r))
    assert k.growth_adjustment_1 == pytest.approx(0.18 / (2 * r) - 0.039 / (2 * r))
    assert k.photosynthesis_arrhenius_temp == pytest.approx(math.log(1.3 / 1.22) / (1 / 285 - 1 / 288))
    assert k.respiration_arrhenius_temp == pytest.approx(math.log(5.429 / 2.785) / (1 / 285 - 1 / 290))
    assert k.maximum_nitrate_uptake == pytest.approx(10 / 0.5 * 24 * 14 / 1e6)
    assert k.maximum_ammonia_uptake == pytest.approx(12 / 0.5 * 24 * 14 / 1e6)
    assert k.temperature_limit.lower_gradient == pytest.approx(1 / 11.8)
    assert k.required_particle_fields() == ("A", "N", "C")
    assert k.coupled_tracers() == ("NO₃", "NH₄", "DIC", "O₂", "DOC", "DON", "bPOC", "bPON")
    p = k.c_params()
    assert p.lower_optimal == 10.0 and p.adapted_latitude == 57.5 and math.isinf(p.exudation_redfield_ratio)


STATES = [  # t, A, N, C, T, NO3, NH4, PAR, u
    (60 * day, 2.0, 1.0, 1.0, 10.0, 10.0, 1.0, 50.0, 0.0),     # the reference test's state (N ≫ N_max)
    (100 * day, 30.0, 0.015, 0.3, 6.0, 4.0, 0.3, 20.0, 0.05),
    (200 * day, 8.0, 0.02, 0.05, 17.5, 0.5, 2.0, 120.0, 0.3),
    (300 * day, 0.5, 0.013, 0.011, 12.0, 12.0, 0.01, 0.5, 0.0),
]


@pytest.mark.parametrize("CN", [math.inf, 12.0])
@pytest.mark.parametrize("state", STATES)
def test_kelp_and_tracer_budgets_close(oracle, state, CN):
    """d/dt of the kelp's nitrogen A kₐ (N + Nₛ) / (14·10⁻³) equals minus the sum of the NO₃, NH₄, DON, bPON tracer
    terms, and the same for carbon with DIC, DOC, bPOC — the identities behind test/test_sugar_kelp.jl:76-89."""
    t, A, N, Cr, T, NO3, NH4, PAR, u = state
    k = ob.SugarKelp(exudation_redfield_ratio=CN)
    p = k.c_params()
    f = lambda name: oracle.kelp(p, name, t, A, N, Cr, T, NO3, NH4, PAR, u=u)  # noqa: E731
    dA, dN, dC = f("A"), f("N"), f("C")
    kA, Ns, Cs = k.structural_dry_weight_per_area, k.structural_nitrogen, k.structural_carbon
    kelp_N = (dA * (N + Ns) + A * dN) * kA / (14 * 0.001)
    kelp_C = (dA * (Cr + Cs) + A * dC) * kA / (12 * 0.001)
    tr_N = [f("NO₃"), f("NH₄"), f("DON"), f("bPON")]
    tr_C = [f("DIC"), f("DOC"), f("bPOC")]
    assert abs(kelp_N + sum(tr_N)) <= 1e-13 * (abs(kelp_N) + sum(abs(x) for x in tr_N))
    assert abs(kelp_C + sum(tr_C)) <= 1e-13 * (abs(kelp_C) + sum(abs(x) for x in tr_C))
    assert f("O₂") == -f("DIC")
    assert f("bPON") > 0 and f("bPOC") > 0 and f("NO₃") <= 0 and f("NH₄") <= 0
    if math.isinf(CN):
        assert f("DON") == 0.0


def test_light_inhibition_solves_its_defining_equation(oracle):
    k = ob.SugarKelp()
    p = k.c_params()
    a, Is = k.photosynthetic_efficiency, k.saturation_irradiance
    for Pm in (0.01, 0.025, 0.031):
        beta, n = oracle.kelp_light_inhibition(p, Pm)
        lhs = a / math.log(1 + a / beta) * (a / (a + beta)) * (beta / (a + beta)) ** (beta / a)
        assert abs(lhs - Pm / Is) <= 1e-15 and beta > 0
        assert 1 <= n <= 1000  # the reference's tolerance eps(1e-9) is rarely met: it usually runs to its cap


def test_seasonal_limitation(oracle):
    p = params()
    vals = [oracle.kelp_seasonal_limitation(p, d * day) for d in range(0, 364)]
    # a₁ (1 ± √|λ|) + a₂ with |λ| ≲ 1 (normalised by the change on day 76): ≈ a₂ … 2 a₁ + a₂, largest while the days
    # lengthen fastest (spring)
    assert min(vals) >= 0.3 - 0.05 and max(vals) <= 0.3 + 2 * 0.85 + 0.05
    assert 60 <= int(np.argmax(vals)) <= 100
    # piecewise constant within a day, periodic over 364 days
    assert oracle.kelp_seasonal_limitation(p, 60 * day) == oracle.kelp_seasonal_limitation(p, 60.9 * day)
    assert oracle.kelp_seasonal_limitation(p, 10 * day) == oracle.kelp_seasonal_limitation(p, 374 * day)
    lib = _lib.load()  # host-side evaluation inside the CUDA library (no device needed)
    for d in (0, 59.5, 171, 300):
        assert lib.obm_kelp_seasonal_limitation(p, d * day) == pytest.approx(oracle.kelp_seasonal_limitation(p, d * day), rel=1e-14)


def test_nearest_node(oracle):
    """3 × 3 × 3 grid of unit cells, z ∈ [−3, 0] (test/test_particles.jl:26): x = y = 0.5, z = 0 deposits in (1, 1, 3)."""
    zf = np.arange(-6.0, 4.0)  # Hz = 3 halo faces either side
    g = oracle.Grid(3, 3, 3, 3, 3, 3, 0.5 * (zf[:-1] + zf[1:]), zf)
    P, B = _lib.OBM_TOPO_PERIODIC, _lib.OBM_TOPO_BOUNDED
    x = np.array([0.5, 1.49, 1.51, 2.9, 3.4, -0.2, 0.5])
    y = np.array([0.5, 0.5, 0.5, 2.5, 0.5, 0.5, 2.6])
    z = np.array([0.0, -0.4, -1.6, -2.9, -5.0, 7.0, -1.5])
    one = np.ones_like(x)
    q, keep = oracle.make_particles(x, y, z, one, one, one, None, 0.5, 1.0, 0.5, 1.0, (P, P, B))
    sy, sz = 3 + 6, (3 + 6) * (3 + 6)
    cell = lambda i, j, k: (i - 1 + 3) + sy * (j - 1 + 3) + sz * (k - 1 + 3)  # noqa: E731  1-based (i, j, k)
    want = [cell(1, 1, 3), cell(2, 1, 3), cell(2, 1, 2), cell(3, 3, 1), cell(1, 1, 1), cell(3, 1, 3), cell(1, 3, 2)]
    assert [oracle.particle_cell(g, q, n) for n in range(7)] == want


def test_scatter_and_step_drivers(oracle):
    """update_tracer_tendencies.jl / time_stepping.jl drivers: deposit = scalefactor · kelp(Val(c)) / volume into the
    nearest cell (particles sharing a cell add up), Euler step from the un-stepped state."""
    zf = np.linspace(-8.0, 4.0, 13)  # Nz = 6, Hz = 3, Δz = 1
    g = oracle.Grid(4, 2, 6, 3, 3, 3, 0.5 * (zf[:-1] + zf[1:]), zf)
    rng = np.random.default_rng(5)
    fld = lambda lo, hi: rng.uniform(lo, hi, size=g.parent_shape)  # noqa: E731
    T, NO3, NH4, PAR = fld(5, 15), fld(1, 10), fld(0.1, 2), fld(5, 80)
    f, keep_f = oracle.make_kelp_tracers(T, NO3, NH4, PAR)
    n = 5
    x, y, z = np.array([0.6, 0.7, 2.2, 3.9, 1.1]), np.array([0.3, 0.4, 1.7, 0.2, 1.2]), np.array([-0.2, -0.3, -2.6, -5.9, -3.3])
    A, N, Cr = rng.uniform(1, 20, n), rng.uniform(0.013, 0.021, n), rng.uniform(0.05, 0.5, n)
    sf = np.array([1.0, 0.5, 2.0, 1.0, 3.0])
    P, B = _lib.OBM_TOPO_PERIODIC, _lib.OBM_TOPO_BOUNDED
    dx, dy = 1.0, 2.0
    q, keep = oracle.make_particles(x, y, z, A.copy(), N.copy(), Cr.copy(), sf, 0.5, dx, 1.0, dy, (P, P, B))
    p = params()
    t = 90 * day
    G = [np.zeros(g.parent_shape) for _ in range(8)]
    G[5] = None  # DON not coupled in this model
    oracle.kelp_update_tendencies(g, p, q, f, G, t)
    cells = [oracle.particle_cell(g, q, i) for i in range(n)]
    assert cells[0] == cells[1]  # two particles in one cell
    for c, name in enumerate(ob.SugarKelp().coupled_tracers()):
        if G[c] is None:
            continue
        want = np.zeros(g.parent_shape).ravel()
        for i in range(n):
            idx = cells[i]
            v = oracle.kelp(p, name, t, A[i], N[i], Cr[i], T.ravel()[idx], NO3.ravel()[idx], NH4.ravel()[idx], PAR.ravel()[idx])
            want[idx] += sf[i] * v / (dx * dy * 1.0)
        np.testing.assert_allclose(G[c].ravel(), want, rtol=1e-15, atol=0)
    out = [np.zeros(n) for _ in range(3)]
    oracle.kelp_step(g, p, q, f, t, 30.0, out)
    for j, name in enumerate(("A", "N", "C")):
        d = np.array([oracle.kelp(p, name, t, A[i], N[i], Cr[i], T.ravel()[cells[i]], NO3.ravel()[cells[i]], NH4.ravel()[cells[i]],
                                  PAR.ravel()[cells[i]]) for i in range(n)])
        assert np.array_equal(out[j], d)
        assert np.array_equal(keep[3 + j], (A, N, Cr)[j] + d * 30.0)


# ---- independent restatement ------------------------------------------------------------------------------------------

def _golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "kelp_rates.json")) as fh:
        return json.load(fh)


def test_golden_vectors_are_what_the_restatement_produces():
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import pyref_kelp as ref
    g = _golden()
    assert tuple(g["names"]) == ref.NAMES and len(g["cases"]) == 36
    for case in g["cases"][::5]:
        kelp, s = ref.SugarKelp(**case["parameters"]), case["state"]
        rates = [kelp(n, s["t"], s["A"], s["N"], s["C"], s["u"], s["v"], s["w"], s["T"], s["NO3"], s["NH4"], s["PAR"])
                 for n in ref.NAMES]
        assert rates == case["rates"]


def test_c_oracle_matches_independent_restatement(oracle):
    """All eleven rates at 12 states × 3 parameter sets, relative 1e-13 (measured: identical to the last bit)."""
    g = _golden()
    assert tuple(g["names"]) == tuple(oracle.KELP_NAMES)
    worst = 0.0
    for case in g["cases"]:
        p, s = ob.SugarKelp(**case["parameters"]).c_params(), case["state"]
        for name, want in zip(g["names"], case["rates"]):
            got = oracle.kelp(p, name, s["t"], s["A"], s["N"], s["C"], s["T"], s["NO3"], s["NH4"], s["PAR"], s["u"], s["v"], s["w"])
            worst = max(worst, abs(got - want) / abs(want) if want else abs(got))
    assert worst <= 1e-13, worst


def test_host_parameter_defaults_match_the_restatement():
    """Every keyword default of SugarKelp.jl:88-139 as the host mirror holds it ≡ as the restatement derives it."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import pyref_kelp as ref
    mine, theirs = ob.SugarKelp(), ref.SugarKelp()
    checked = 0
    for name, want in vars(theirs).items():
        if name == "temperature_limit":
            for f in ("lower_optimal", "upper_optimal", "lower_gradient", "upper_gradient"):
                assert getattr(mine.temperature_limit, f) == pytest.approx(getattr(want, f), rel=1e-15)
            continue
        assert getattr(mine, name) == pytest.approx(want, rel=1e-15), name
        checked += 1
    assert checked >= 40
