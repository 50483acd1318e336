"""Pin the ORACLE's NPD family on the properties the reference's own tests assert
(test/test_NutrientsPlanktonDetritus.jl:41-139): the all-zero state has exactly zero tendencies,
and total nitrogen / carbon (with scale factors) is conserved for all 36 non-sinking variants —
both instantaneously (Σ tendencies = 0) and over 100 unit time steps (rtol √eps)."""
import itertools

import numpy as np
import pytest

import oceanbiome_b200 as ob

NUTRIENTS = (ob.NitrateAmmonia, ob.NitrateAmmoniaIron, ob.Nutrient)
DETRITUS = (ob.TwoParticleAndDissolved, ob.VariableRedfieldDetritus, ob.Detritus)
VARIANTS = list(itertools.product(DETRITUS, (None, ob.Oxygen), (None, ob.CarbonateSystem), NUTRIENTS))


def make(det, oxy, car, nut):
    return ob.NutrientsPlanktonDetritus(nutrients=nut(), plankton=ob.PhytoZoo(), detritus=det(),
                                        carbonate_system=car() if car else None, oxygen=oxy() if oxy else None)


def random_state(bgc, rng):
    """initial values as in test_NutrientsPlanktonDetritus.jl:57-91"""
    v = {"P": rng.random(), "Z": rng.random(), "NO₃": 10 * rng.random(), "NH₄": rng.random(), "Fe": 1e-3 * rng.random(),
         "N": 10 * rng.random(), "DIC": 2000 * rng.random(), "Alk": 2000 * rng.random(), "O₂": 300 * rng.random(),
         "sPOM": rng.random(), "bPOM": rng.random(), "DOM": rng.random(), "D": rng.random(), "T": 10.0}
    v["sPON"], v["bPON"], v["DON"] = rng.random(), rng.random(), rng.random()
    v["sPOC"], v["bPOC"], v["DOC"] = 6.56 * v["sPON"], 6.56 * v["bPON"], 6.56 * v["DON"]
    return {n: v[n] for n in bgc.required_biogeochemical_tracers()}


def tendencies(oracle, og, bgc, state, PAR=100.0):
    names = bgc.required_biogeochemical_tracers()
    tr = []
    for n in names:
        a = np.zeros(og.parent_shape)
        og.interior(a)[...] = state[n]
        tr.append(a)
    par = np.full(og.parent_shape, PAR)
    G = oracle.npd_tendencies(og, bgc.c_params(), tr, par)
    return {n: float(og.interior(g)[0, 0, 0]) for n, g in zip(names, G)}


def totals(bgc, values):
    cons = bgc.conserved_tracers(labeled=True)
    out = {"nitrogen": sum(values[n] for n in cons["nitrogen"])}
    if "carbon" in cons:
        out["carbon"] = sum(values[n] * f for n, f in zip(cons["carbon"]["tracers"], cons["carbon"]["scalefactors"]))
    return out


@pytest.fixture(scope="module")
def one(oracle):
    g = ob.RectilinearGrid(size=(1, 1, 1), extent=(1, 1, 2), device="cpu")
    return oracle.Grid.like(g)


@pytest.mark.parametrize("det,oxy,car,nut", VARIANTS)
def test_names_match_reference_order(oracle, det, oxy, car, nut):
    bgc = make(det, oxy, car, nut)
    assert oracle.npd_tracer_names(bgc.c_params()) == bgc.required_biogeochemical_tracers()


@pytest.mark.parametrize("det,oxy,car,nut", VARIANTS)
def test_zero_state_has_zero_tendencies(oracle, one, det, oxy, car, nut):
    # test_NutrientsPlanktonDetritus.jl:101-112 — this is what the eps(0.0) guards are for (no 0/0 NaN)
    bgc = make(det, oxy, car, nut)
    G = tendencies(oracle, one, bgc, {n: 0.0 for n in bgc.required_biogeochemical_tracers()})
    assert all(v == 0.0 for v in G.values()), G


@pytest.mark.parametrize("det,oxy,car,nut", VARIANTS)
def test_instantaneous_conservation(oracle, one, det, oxy, car, nut):
    bgc = make(det, oxy, car, nut)
    rng = np.random.default_rng(42)
    for _ in range(5):
        G = tendencies(oracle, one, bgc, random_state(bgc, rng))
        cons = bgc.conserved_tracers(labeled=True)
        sN = sum(G[n] for n in cons["nitrogen"])
        aN = sum(abs(G[n]) for n in cons["nitrogen"])
        assert abs(sN) <= 4e-16 * aN
        if "carbon" in cons:
            terms = [G[n] * f for n, f in zip(cons["carbon"]["tracers"], cons["carbon"]["scalefactors"])]
            assert abs(sum(terms)) <= 4e-16 * sum(abs(t) for t in terms)


@pytest.mark.parametrize("det,oxy,car,nut", VARIANTS[::5])
def test_conservation_over_100_steps(oracle, one, det, oxy, car, nut):
    # test_NutrientsPlanktonDetritus.jl:114-139 (forward Euler here; Δt = 1 s as in the reference)
    bgc = make(det, oxy, car, nut)
    state = random_state(bgc, np.random.default_rng(42))
    t0 = totals(bgc, state)
    for _ in range(100):
        G = tendencies(oracle, one, bgc, state)
        state = {n: state[n] + 1.0 * G[n] for n in state}
    t1 = totals(bgc, state)
    for k in t0:
        assert abs(t1[k] - t0[k]) <= 1.5e-8 * abs(t0[k])


def test_npzd_defaults_and_oxygen_dead_dispatch(oracle, one):
    """NPZD(grid) tracer order (:N, :P, :Z, :T, :D) constructors.jl:169; with Oxygen the `<:Nutrient`
    specialisation is unreachable, so ∂ₜO₂ = Rp μP exactly (SURVEY App. A bug 2)."""
    g = ob.RectilinearGrid(size=(1, 1, 1), extent=(1, 1, 2), device="cpu")
    bgc = ob.NPZD(g, oxygen=ob.Oxygen()).underlying_biogeochemistry
    assert bgc.required_biogeochemical_tracers() == ("N", "P", "Z", "T", "D", "O₂")
    state = {"N": 3.0, "P": 0.4, "Z": 0.2, "T": 12.0, "D": 0.3, "O₂": 200.0}
    G = tendencies(oracle, one, bgc, state, PAR=60.0)
    p = bgc.plankton
    kPAR = p.light_half_saturation
    mu = (p.phytoplankton_maximum_growth_rate * (60.0 / np.sqrt(60.0 ** 2 + kPAR ** 2))
          * (3.0 / (3.0 + p.nitrate_half_saturation)) * 1.88 ** (12.0 / 10) * 0.4)
    assert np.isclose(G["O₂"], 10.75 * mu, rtol=1e-14)
    assert G["T"] == 0.0
    assert abs(G["N"] + G["P"] + G["Z"] + G["D"]) <= 4e-16 * sum(abs(G[n]) for n in "NPZD")


# ---- absolute values: the C oracle against an independent transliteration of the reference ------------------------
def _golden():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "npd_tendencies.json"), encoding="utf-8"))["cases"]


GOLDEN_MODELS = {
    "lobster": lambda: ob.NutrientsPlanktonDetritus(ob.NitrateAmmonia(), ob.PhytoZoo(), ob.TwoParticleAndDissolved()),
    "lobster_carbonate_oxygen": lambda: ob.NutrientsPlanktonDetritus(ob.NitrateAmmonia(), ob.PhytoZoo(), ob.TwoParticleAndDissolved(),
                                                                     ob.CarbonateSystem(), ob.Oxygen()),
    "lobster_iron_variable_redfield_carbonate_oxygen": lambda: ob.NutrientsPlanktonDetritus(
        ob.NitrateAmmoniaIron(), ob.PhytoZoo(), ob.VariableRedfieldDetritus(), ob.CarbonateSystem(), ob.Oxygen()),
    "npzd": lambda: ob.NPZD(ob.RectilinearGrid(size=(1, 1, 1), extent=(1, 1, 2), device="cpu")).underlying_biogeochemistry,
    "npzd_carbonate_oxygen": lambda: ob.NPZD(ob.RectilinearGrid(size=(1, 1, 1), extent=(1, 1, 2), device="cpu"),
                                             carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen()).underlying_biogeochemistry,
}


@pytest.mark.parametrize("case", sorted(GOLDEN_MODELS))
def test_c_oracle_matches_independent_restatement(oracle, one, case):
    """tests/golden/npd_tendencies.json holds tendencies computed by oracle/pyref_npd.py, a method-by-method Python
    transliteration of the reference that shares no code with oracle_npd.c (scripts/make_npd_golden.py).  Two
    independent readings of the same source must agree to rounding: 1e-15 of the largest term of the model."""
    g = _golden()[case]
    bgc = GOLDEN_MODELS[case]()
    assert list(bgc.required_biogeochemical_tracers()) == g["tracers"]
    for row in g["rows"]:
        got = tendencies(oracle, one, bgc, row["state"], PAR=row["state"]["PAR"])
        scale = max(abs(v) for v in row["tendencies"].values())
        for n, want in row["tendencies"].items():
            assert abs(got[n] - want) <= 2e-15 * scale, (case, n, got[n], want)


@pytest.mark.parametrize("det,oxy,car,nut", VARIANTS[::5])
def test_term_scales_bound_the_tendencies(oracle, one, det, oxy, car, nut):
    """`orc_npd_tendency_scales` — the S of the parity metric |a − b| ≤ 1e-12·max(|b|, S): Σ|additive terms| of
    each tendency bounds the tendency itself, and is 0 exactly where the model gives the tracer no method (`zero(grid)`)."""
    bgc = make(det, oxy, car, nut)
    rng = np.random.default_rng(3)
    names = bgc.required_biogeochemical_tracers()
    for _ in range(20):
        state = random_state(bgc, rng)
        tr = []
        for n in names:
            a = np.zeros(one.parent_shape)
            one.interior(a)[...] = state[n]
            tr.append(a)
        par = np.full(one.parent_shape, 100.0 * rng.random())
        G = oracle.npd_tendencies(one, bgc.c_params(), tr, par)
        S = oracle.npd_tendency_scales(one, bgc.c_params(), tr, par)
        for n, g, s in zip(names, G, S):
            t, sc = float(one.interior(g)[0, 0, 0]), float(one.interior(s)[0, 0, 0])
            assert sc >= 0 and abs(t) <= sc * (1 + 1e-14), (n, t, sc)
            if n == "T":
                assert sc == 0 and t == 0
