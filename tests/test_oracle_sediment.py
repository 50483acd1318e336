"""ORACLE checks for the sediment path (CPU only).  The reference's own sediment testset is commented out
(test/test_sediments.jl:106-163) — PARITY UNPINNED — so the oracle is checked on what is derivable from the
equations: the burial-efficiency closed form, nitrogen bookkeeping of both models (what sinks in is stored,
buried or returned), the ifelse(isfinite) guards, and the integer bottom-index search."""
import math

import numpy as np
import pytest

import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib as abi


def params(model, grid=None, **kw):
    g = grid or ob.RectilinearGrid(size=(2, 2, 4), extent=(2, 2, 400), device="cpu")
    return model(g, **kw).c_params()


def test_instant_remineralisation_closed_form(oracle):
    p = params(ob.InstantRemineralisationSediment)
    for flux in (0.0, 1e-6, 0.3, 5.0):
        (dS,), (ret,) = oracle.sediment_point(p, [0.0], 0, 0, 0, flux)
        e = 0.013 + 0.53 * (flux / (7.0 / 6.56 + flux)) ** 2
        assert math.isclose(dS, e * flux, rel_tol=1e-15, abs_tol=0) and math.isclose(ret, (1 - e) * flux, rel_tol=1e-15, abs_tol=0)
        assert math.isclose(dS + ret, flux, rel_tol=1e-15, abs_tol=0)  # everything that sinks is buried or returned


def test_simple_multi_g_nitrogen_bookkeeping(oracle):
    """Σ pool tendencies = sinking N − (λs Ns + λf Nf), and the remineralised N returns as NO₃ + NH₄
    (simple_multi_G.jl:165-226): pₙ Nr − 0.8 pₙ′ Cr + (1 − pₙ) Nr + 0.8 pₙ′ Cr = Nr."""
    p = params(ob.SimpleMultiGSediment)
    rng = np.random.default_rng(5)
    for _ in range(20):
        pools = list(10 ** rng.uniform(-2, 1, 3))
        NO3, NH4, O2, fN = rng.uniform(1, 30), rng.uniform(0.1, 5), rng.uniform(50, 400), 10 ** rng.uniform(-7, -4)
        dP, (fNO3, fNH4, fO2) = oracle.sediment_point(p, pools, NO3, NH4, O2, fN)
        Nr = p.slow_decay_rate * pools[0] + p.fast_decay_rate * pools[1]
        assert math.isclose(sum(dP), fN - Nr, rel_tol=1e-12)
        assert math.isclose(fNO3 + fNH4, Nr, rel_tol=1e-12)
        assert fO2 < 0  # the sediment consumes oxygen


def test_simple_multi_g_carbon_variant(oracle):
    p = params(ob.SimpleMultiGSediment, sinking_carbon=("sPOC", "bPOC"), sinking_nitrogen=("sPON", "bPON"))
    assert p.carbon == 1 and p.nsinking_carbon == 2
    pools = [0.5, 0.2, 0.1, 3.0, 1.5, 0.6]
    dP, cf = oracle.sediment_point(p, pools, 10.0, 1.0, 200.0, 2e-6, 1.3e-5)
    assert len(dP) == 6 and len(cf) == 4
    Cr = p.slow_decay_rate * pools[3] + p.fast_decay_rate * pools[4]
    assert math.isclose(cf[3], Cr, rel_tol=1e-15)                       # DIC (simple_multi_G.jl:361-368)
    assert math.isclose(sum(dP[3:]), 1.3e-5 - Cr, rel_tol=1e-12)        # carbon bookkeeping


def test_empty_sediment_guards(oracle):
    """Empty pools: Cr = 0 ⇒ log(0) = −Inf ⇒ p non-finite ⇒ ifelse(isfinite(p), p, 0) (simple_multi_G.jl:394,410,424)."""
    p = params(ob.SimpleMultiGSediment)
    dP, cf = oracle.sediment_point(p, [0.0, 0.0, 0.0], 10.0, 1.0, 200.0, 0.0)
    assert dP == [0.0, 0.0, 0.0] and all(v == 0.0 for v in cf)


def test_bottom_indices_bit_exact(oracle):
    g = ob.RectilinearGrid(size=(6, 5, 10), extent=(6, 5, 100), device="cpu")
    og = oracle.Grid.like(g)
    h = np.zeros(og.plane_shape)
    rng = np.random.default_rng(0)
    og.interior(h)[...] = rng.uniform(-120, 10, (1, 5, 6))
    kb = oracle.find_bottom_cells(og, h)
    zc = g.zc
    for j in range(5):
        for i in range(6):
            hh = og.interior(h)[0, j, i]
            want = 1
            while zc[want - 1] <= hh and want < 10:
                want += 1
            assert og.interior(kb)[0, j, i] == want


def test_constructors(oracle):
    g = ob.RectilinearGrid(size=(3, 3, 30), extent=(10, 10, 200), device="cpu")
    sed = ob.SimpleMultiGSediment(g)
    assert tuple(sed.fields) == ("Ns", "Nf", "Nr") and tuple(sed.tracked_fields) == ("NO₃", "NH₄", "O₂", "sPOM", "bPOM")
    assert math.isclose(sed.biogeochemistry.sedimentation_rate, 982 * abs(g.zc[0]) ** -1.548)
    ir = ob.InstantRemineralisationSediment(g, sinking_tracers=("sPOM", "bPOM"), remineralisation_reciever="NH₄")
    assert tuple(ir.fields) == ("storage",) and ir.biogeochemistry.coupled_tracers() == ("NH₄",)
    with pytest.raises(ValueError, match="not configured for sediment models"):
        ob.BiogeochemicalSediment(g, ob.InstantRemineralisation(), timestepper="Euler")
    bgc = ob.LOBSTER(g, sediment=ir)
    assert "instant remineralisation" in repr(bgc)


@pytest.mark.parametrize("carbon", [False, True])
def test_c_oracle_matches_independent_restatement(oracle, carbon):
    """oracle/pyref_sediment.py transliterates simple_multi_G.jl method by method (no code shared with
    oracle_sediment.c, reference default parameters typed in again): pool tendencies and the NO₃ / NH₄ / O₂ / DIC return
    fluxes agree to rounding at seeded states, including the guard for an exactly empty sediment."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pyref_sediment as ref
    g = ob.RectilinearGrid(size=(2, 2, 4), extent=(2, 2, 400), device="cpu")
    kw = dict(sinking_carbon=("sPOC", "bPOC"), sinking_nitrogen=("sPON", "bPON")) if carbon else {}
    p = params(ob.SimpleMultiGSediment, grid=g, **kw)
    model = ref.SimpleMultiG(float(g.zc[0]), carbon=carbon)
    rng = np.random.default_rng(17)
    names = ("Ns", "Nf", "Nr") + (("Cs", "Cf", "Cr") if carbon else ())
    coupled = ("NO₃", "NH₄", "O₂") + (("DIC",) if carbon else ())
    for case in range(12):
        pools = list(10 ** rng.uniform(-3, 1, len(names)))
        NO3, NH4, O2 = rng.uniform(0.5, 30), rng.uniform(0.05, 5), rng.uniform(2, 400)
        fN, fC = 10 ** rng.uniform(-8, -4), 10 ** rng.uniform(-7, -3)
        if case == 0:
            pools = [0.0] * len(names)  # log(0) ⇒ non-finite fractions ⇒ ifelse(isfinite(p), p, 0)
        dP, cf = oracle.sediment_point(p, pools, NO3, NH4, O2, fN, *([fC] if carbon else []))
        for n, got in zip(names, dP):
            want = model(n, pools, NO3, NH4, O2, fN, fC)
            assert math.isclose(got, want, rel_tol=1e-14, abs_tol=1e-300), (case, n, got, want)
        for n, got in zip(coupled, cf):
            want = model(n, pools, NO3, NH4, O2, fN, fC)
            assert math.isclose(got, want, rel_tol=1e-11, abs_tol=1e-300), (case, n, got, want)  # exp of ≈ 10 log-products
