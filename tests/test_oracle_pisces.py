"""Pin the ORACLE's PISCES on what the reference's own test asserts (test/test_PISCES.jl:32-92): at
PISCES_INITIAL_VALUES (box model, z = −5, PAR₁₂₃ = 100, PAR = 300, zₘₓₗ = zₑᵤ = −10, κ̄ = 1, PAR̄ = 300, w = 0,
iron scavenging enhancement and N-fixation off) the carbon, iron, silicon, phosphate and nitrogen budgets of
the 24 tendencies close to atol 1e-20 … 1e-22, and the all-zero state has exactly zero tendencies.
Cross-check: the 24 smoke values derived independently in SURVEY §8c (Ω = 3, t = 1.6 s)."""
import math

import numpy as np
import pytest

import oceanbiome_b200 as ob
from oceanbiome_b200 import pisces

SURVEY_SMOKE = {  # SURVEY.md §8c "Readings already validated", mmol m⁻³ s⁻¹ (Fe in μmol)
    "P": 1.683697e-06, "PChl": 9.841187e-08, "PFe": 6.841848e-08, "D": 1.121333e-07, "DChl": 5.410023e-09,
    "DFe": 4.426870e-09, "DSi": 4.689823e-08, "Z": 3.130033e-08, "M": 4.732613e-08, "DOC": 3.978079e-07,
    "POC": 1.017150e-06, "GOC": -1.586144e-06, "SFe": 8.212952e-08, "BFe": -8.313395e-08, "PSi": -2.725922e-08,
    "CaCO₃": 3.631553e-09, "NO₃": -4.978455e-08, "NH₄": -1.735952e-07, "PO₄": -1.396123e-08, "Fe": -7.286382e-08,
    "Si": -1.963901e-08, "DIC": -1.706902e-06, "Alk": -1.310737e-07, "O₂": 1.956413e-06}


@pytest.fixture(scope="module")
def box():
    g = ob.RectilinearGrid(size=(1,), z=(-10, 0), topology=("Flat", "Flat", "Bounded"), device="cpu")
    bgc = ob.PISCES(g, sinking_speeds={"POC": 0.0, "GOC": 0.0}, iron=pisces.SimpleIron(excess_scavenging_enhancement=0.0),
                    nitrogen=pisces.NitrateAmmonia(maximum_fixation_rate=0.0))
    return bgc.underlying_biogeochemistry


def point(oracle, u, values, t=1.0, Omega=0.0):
    G = oracle.pisces_point(u.c_params(t), [values[n] for n in pisces.TRACERS], 100.0, 100.0, 100.0, 300.0, Omega, 0.0, 0.0,
                            -10.0, -10.0, 1.0, 300.0, -5.0)
    return dict(zip(pisces.TRACERS, G))


def test_tracer_and_auxiliary_lists(box):
    assert box.required_biogeochemical_tracers() == (
        "P", "PChl", "PFe", "D", "DChl", "DFe", "DSi", "Z", "M", "DOC", "POC", "GOC", "SFe", "BFe", "PSi", "CaCO₃", "NO₃",
        "NH₄", "PO₄", "Fe", "Si", "DIC", "Alk", "O₂", "T", "S")
    assert box.required_biogeochemical_auxiliary_fields() == (
        "zₘₓₗ", "zₑᵤ", "Si′", "Ω", "κ", "mixed_layer_PAR", "wPOC", "wGOC", "PAR", "PAR₁", "PAR₂", "PAR₃")


def test_zero_state_is_exactly_zero(oracle, box):
    # test_PISCES.jl:66-69
    G = point(oracle, box, {n: 0.0 for n in pisces.TRACERS})
    assert all(v == 0.0 for v in G.values()), {k: v for k, v in G.items() if v != 0}


@pytest.mark.parametrize("Omega,t", [(0.0, 1.0), (3.0, 1.6), (0.7, 86400.0 * 200)])
def test_element_budgets_close(oracle, box, Omega, t):
    # test_PISCES.jl:71-90 with the reference's own tolerances
    G = point(oracle, box, pisces.PISCES_INITIAL_VALUES, t=t, Omega=Omega)
    cons = box.conserved_tracers(ntuple=True)
    assert abs(sum(G[n] for n in cons["carbon"])) <= 1e-20
    assert abs(sum(G[n] * f for n, f in zip(cons["iron"]["tracers"], cons["iron"]["scalefactors"]))) <= 1e-21
    assert abs(sum(G[n] for n in cons["silicon"])) <= 1e-21
    assert abs(sum(G[n] * f for n, f in zip(cons["phosphate"]["tracers"], cons["phosphate"]["scalefactors"]))) <= 1e-22
    assert abs(sum(G[n] * f for n, f in zip(cons["nitrogen"]["tracers"], cons["nitrogen"]["scalefactors"]))) <= 1e-21
    assert G["T"] == 0.0 and G["S"] == 0.0


def test_agrees_with_independently_derived_smoke_values(oracle, box):
    G = point(oracle, box, pisces.PISCES_INITIAL_VALUES, t=1.6, Omega=3.0)
    for n, v in SURVEY_SMOKE.items():
        assert math.isclose(G[n], v, rel_tol=1e-6), (n, G[n], v)


def test_budgets_close_on_random_states(oracle, box):
    rng = np.random.default_rng(11)
    cons = box.conserved_tracers(ntuple=True)
    for _ in range(50):
        vals = {n: v * math.exp(rng.uniform(-1, 1)) for n, v in pisces.PISCES_INITIAL_VALUES.items()}
        vals["T"] = rng.uniform(0, 30)
        G = oracle.pisces_point(box.c_params(rng.uniform(0, 3e7)), [vals[n] for n in pisces.TRACERS], *rng.uniform(0, 80, 3), 90.0,
                                rng.uniform(0.3, 4), -2 / 86400, -40 / 86400, rng.uniform(-100, -5), rng.uniform(-100, -5),
                                10 ** rng.uniform(-4, 0), rng.uniform(0, 100), rng.uniform(-200, -1))
        G = dict(zip(pisces.TRACERS, G))
        for key in ("carbon", "silicon"):
            terms = [G[n] for n in cons[key]]
            assert abs(sum(terms)) <= 1e-15 * sum(abs(x) for x in terms)
        for key in ("iron", "phosphate", "nitrogen"):
            terms = [G[n] * f for n, f in zip(cons[key]["tracers"], cons[key]["scalefactors"])]
            assert abs(sum(terms)) <= 1e-15 * sum(abs(x) for x in terms)


def test_day_length_call_orders(oracle, box):
    """The reference calls day_length(φ, t) in growth rates and day_length(t, φ) in chlorophyll synthesis
    (growth_rate.jl:30 vs :143); both must differ and match the C restatement of Utils.jl:13-34."""
    for t in (0.0, 1.6, 86400.0 * 100.5):
        p = box.c_params(t)
        assert math.isclose(p.day_length_growth, oracle.cbm_day_length(45.0, t), rel_tol=1e-14)
        assert math.isclose(p.day_length_chlorophyll, oracle.cbm_day_length(t, 45.0), rel_tol=1e-14)
        assert abs(p.day_length_growth - p.day_length_chlorophyll) > 1000
    # Forsythe et al. 1995: ~12 h at the equinoxes, longer in summer at 45°N
    assert abs(oracle.cbm_day_length(86400.0 * 80, 45.0) / 3600 - 12.2) < 0.3
    assert oracle.cbm_day_length(86400.0 * 172, 45.0) / 3600 > 15


def test_conserved_groups_and_scalers(box):
    g = ob.RectilinearGrid(size=(1,), z=(-10, 0), topology=("Flat", "Flat", "Bounded"), device="cpu")
    bgc = ob.PISCES(g, scale_negatives=True)
    assert len(bgc.modifiers) == 5
    assert bgc.modifiers[1].tracers == ("PFe", "DFe", "Z", "M", "SFe", "BFe", "Fe")
    assert bgc.modifiers[1].scalefactors == (1, 1, 0.01, 0.015, 1, 1, 1)
    assert bgc.modifiers[3].tracers == ("DSi", "Si", "PSi")
    assert list(bgc.biogeochemical_auxiliary_fields())[:4] == ["zₘₓₗ", "zₑᵤ", "Si′", "Ω"]
    # wPOC: faces 1…Nz = −2/day, top face 0 (sinking_velocity_fields.jl:15-17)
    w = bgc.underlying_biogeochemistry.sinking_velocities["POC"].face_interior[:, 0, 0]
    assert float(w[0]) == -2 / 86400 and float(w[-1]) == 0.0


# ---- absolute values: the C oracle against an independent transliteration of the reference ------------------------
def _golden_rows():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pisces_tendencies.json"), encoding="utf-8"))["rows"]


@pytest.mark.parametrize("row", range(14))
def test_c_oracle_matches_independent_restatement(oracle, row):
    """tests/golden/pisces_tendencies.json holds the 24 tendencies computed by oracle/pyref_pisces.py — a
    method-by-method Python transliteration of the reference that shares no code with oracle_pisces.c
    (scripts/make_pisces_golden.py) — at 14 states that take both sides of every branch.  Two independent readings of
    the same source agree to rounding: 1e-13 of the largest un-cancelled flux of the cell."""
    r = _golden_rows()[row]
    f = r["state"]
    g = ob.RectilinearGrid(size=(1,), z=(-10, 0), topology=("Flat", "Flat", "Bounded"), device="cpu")
    u = ob.PISCES(g, latitude=pisces.PrescribedLatitude(r["latitude"])).underlying_biogeochemistry
    p = u.c_params(f["t"])
    # the host-evaluated day lengths (both argument orders of the reference) agree with the transliteration's
    assert math.isclose(p.day_length_growth, r["day_length_growth"], rel_tol=1e-13)
    assert math.isclose(p.day_length_chlorophyll, r["day_length_chlorophyll"], rel_tol=1e-13)
    vals = [f.get(n, 0.0) for n in pisces.TRACERS]
    G = dict(zip(pisces.TRACERS, oracle.pisces_point(p, vals, f["PAR₁"], f["PAR₂"], f["PAR₃"], f["PAR"], f["Ω"], f["wPOC"], f["wGOC"],
                                                     f["zₘₓₗ"], f["zₑᵤ"], f["κ"], f["mixed_layer_PAR"], f["z"])))
    want = r["tendencies"]
    carbon = max(abs(want[n]) for n in ("P", "D", "Z", "M", "DOC", "POC", "GOC", "DIC"))
    for n, v in want.items():
        scale = max(abs(v), 1e-3 * carbon if n not in ("PChl", "DChl", "PFe", "DFe", "SFe", "BFe", "Fe") else 0.0)
        assert abs(G[n] - v) <= 1e-13 * max(scale, 1e-300) + 1e-30, (row, n, G[n], v)


# ---- what the kernels' fast pass may and may not do with min / max --------------------------------------------------

def _random_point(rng, extreme):
    """A state with finite inputs: ordinary (factor e^{±2} around the reference test's state) or extreme (zeros,
    denormals, huge values — what a blown-up simulation hands over just before it produces its first NaN)."""
    if not extreme:
        vals = {n: v * math.exp(rng.uniform(-2, 2)) for n, v in pisces.PISCES_INITIAL_VALUES.items()}
        vals["T"], vals["S"] = rng.uniform(-2, 32), rng.uniform(20, 40)
    else:
        pick = lambda: rng.choice([0.0, 5e-324, 1e-300, 1e-30, 1e-6, 1.0, 1e6, 1e30, 1e150])  # noqa: E731
        vals = {n: pick() for n in pisces.PISCES_INITIAL_VALUES}
        for biomass in ("P", "D", "Z", "M"):  # incl. positive biomasses far below their pigment (quota overflow, see below)
            vals[biomass] = rng.choice([0.0, 5e-324, 1e-310, 1e-280, 1e-200, 1e-30, 1e-6, 1.0, 1e6])
        vals["T"], vals["S"] = rng.choice([-2.0, 0.0, 40.0, 1e3]), rng.choice([0.0, 35.0, 1e3])
    aux = [rng.uniform(0, 80), rng.uniform(0, 80), rng.uniform(0, 80), 0.0, rng.choice([0.0, 0.5, 1.0, 3.0, 1e6]),
           -rng.choice([0.0, 2.0, 50.0]) / 86400, -rng.choice([1e-9, 30.0, 200.0]) / 86400, -rng.uniform(1, 300),
           -rng.uniform(1, 300), rng.choice([1e-12, 1e-4, 1.0, 1e6]), rng.uniform(0, 300), -rng.uniform(0.5, 400)]
    aux[3] = aux[0] + aux[1] + aux[2]
    return vals, aux


def _lost_nans(oracle, box, vals, aux, t):
    """Tendencies that are finite under select semantics but differ from the reference's (NaN there)."""
    p = box.c_params(t)
    v = [vals.get(n, 0.0) for n in pisces.TRACERS]
    exact = oracle.pisces_point(p, v, *aux)
    fast = oracle.pisces_point(p, v, *aux, select=True)
    return [pisces.TRACERS[n] for n in range(24) if math.isfinite(fast[n]) and not (exact[n] == fast[n])]


@pytest.mark.parametrize("extreme", [False, True])
def test_select_min_max_equals_propagating_min_max_for_finite_inputs(oracle, box, extreme):
    """The kernels' fast pass takes min / max by compare + select, which swallows a NaN operand; cells with a non-finite
    INPUT are routed to the exact pass.  What remains to show is that no NaN is born mid-way from finite inputs and then
    meets ONLY selects.  The oracle rebuilt with select semantics (-DORC_SELECT_MINMAX) is the fast pass's arithmetic:
    over ordinary and extreme finite states, wherever its result is finite it is the reference's bit for bit — a
    non-finite result is recomputed by the exact pass anyway."""
    rng = np.random.default_rng(20261017 + extreme)
    for _ in range(1500):
        vals, aux = _random_point(rng, extreme)
        lost = _lost_nans(oracle, box, vals, aux, rng.uniform(0, 3e7))
        assert not lost, (lost, vals, aux)


def test_quota_overflow_reaches_the_exact_pass(oracle, box):
    """A POSITIVE plankton biomass far below its pigment (here a denormal) makes the iron and chlorophyll quotas overflow:
    θ = Chl / (12 I + eps(0)) = Inf, and L_Fe = min(1, max(0, (θFe − θFe_min) / θopt)) = Inf − Inf = NaN — a NaN born
    mid-way from finite inputs, which the reference's `min` / `max` propagate into a dozen tendencies.  With the
    reference's operand order min(L_N, L_PO₄, L_Fe, L_Si) a compare + select drops it for diatoms (L_Si is finite there);
    the fast pass therefore takes L_Fe last (a select propagates a NaN in its second operand), and the oracle built with
    select semantics does the same: nothing is lost, the cell comes out non-finite and is redone by the exact pass.
    (r01 recorded this case as a known limit of the fast pass.)"""
    vals = dict(pisces.PISCES_INITIAL_VALUES)
    aux = [30.0, 30.0, 30.0, 90.0, 0.8, -2 / 86400, -30 / 86400, -50.0, -80.0, 1e-3, 40.0, -20.0]
    for tiny in (5e-324, 1e-310, 1e-300, 1e-290):
        for who in ("D", "P"):
            v = dict(vals)
            v[who] = tiny
            exact = oracle.pisces_point(box.c_params(1e6), [v.get(n, 0.0) for n in pisces.TRACERS], *aux)
            if tiny == 5e-324:
                assert math.isnan(exact[pisces.TRACERS.index(who)])  # the reference's answer for such a cell
            assert not _lost_nans(oracle, box, v, aux, 1e6), (who, tiny)
    vals["D"] = 0.0  # exactly zero biomass is guarded in the reference itself: finite, nothing to lose
    assert not _lost_nans(oracle, box, vals, aux, 1e6)


def test_select_min_max_swallows_what_the_input_guard_catches(oracle, box):
    """The counter-example that motivates the guard over ALL inputs: Fe = +Inf gives Inf / Inf = NaN inside a min, which a
    select drops — three tendencies come out finite where the reference has NaN."""
    vals = dict(pisces.PISCES_INITIAL_VALUES)
    vals["Fe"] = math.inf
    v = [vals.get(n, 0.0) for n in pisces.TRACERS]
    aux = [30.0, 30.0, 30.0, 90.0, 0.8, -2 / 86400, -30 / 86400, -50.0, -80.0, 1e-3, 40.0, -20.0]
    p = box.c_params(1e6)
    exact = oracle.pisces_point(p, v, *aux)
    fast = oracle.pisces_point(p, v, *aux, select=True)
    lost = [pisces.TRACERS[n] for n in range(24) if math.isnan(exact[n]) and math.isfinite(fast[n])]
    assert lost, "expected the select arithmetic to lose NaNs for an infinite input"


# ---- the parity metric's scale: S = Σ|additive terms| per tendency ------------------------------------

def test_term_scales_bound_the_tendencies_and_compose(oracle, box):
    """`orc_pisces_point_terms`: every tendency is bounded by the sum of the magnitudes of its own additive terms
    (triangle inequality, to rounding); a tendency that is one flux has S = |t|; Alk = NH₄ − NO₃ − 2 CaCO₃
    (inorganic_carbon.jl:49-58) carries S_NH₄ + S_NO₃ + 2 S_CaCO₃; the tendencies are those of `orc_pisces_point` bit for
    bit (recording S does not touch the arithmetic); T and S have no terms."""
    rng = np.random.default_rng(77)
    ix = {n: q for q, n in enumerate(pisces.TRACERS)}
    saw_cancellation = False
    for _ in range(200):
        vals, aux = _random_point(rng, False)
        v = [vals.get(n, 0.0) for n in pisces.TRACERS]
        p = box.c_params(rng.uniform(0, 3e7))
        t, S = oracle.pisces_point_terms(p, v, *aux)
        assert t == oracle.pisces_point(p, v, *aux)
        for n in range(24):
            assert S[n] >= 0 and abs(t[n]) <= S[n] * (1 + 1e-14), (pisces.TRACERS[n], t[n], S[n])
            saw_cancellation |= abs(t[n]) < 1e-2 * S[n]
        assert S[ix["T"]] == 0 and S[ix["S"]] == 0
        assert S[ix["Alk"]] == pytest.approx(S[ix["NH₄"]] + S[ix["NO₃"]] + 2 * S[ix["CaCO₃"]], rel=1e-15)
        # a sum over disjoint positive fluxes: DIC's terms are ≥ each of the fluxes that also close the carbon budget
        assert S[ix["DIC"]] >= abs(t[ix["DIC"]])
    assert saw_cancellation  # the metric matters: some tendencies ARE near-cancelling differences of their terms
