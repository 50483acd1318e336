"""Pin the ORACLE's carbonate chemistry against the reference's own goldens and known-answer tests
(CPU only).  Sources: docstring src/Models/CarbonChemistry/carbon_chemistry.jl:55-62 (run as a
doctest by test/runtests.jl:20-22) and test/test_gasexchange_carbon_chem.jl:113-181."""
import math

import numpy as np
import pytest

from oceanbiome_b200 import _lib as abi


def test_docstring_goldens_bit_exact(oracle):
    # carbon_chemistry.jl:55-62 — all three 16-digit values reproduce to the last bit
    assert oracle.carbon_chemistry(2000.0, 10.0, 35.0, Alk=2000.0) == 1308.1474527899106
    assert oracle.carbon_chemistry(2000.0, 10.0, 35.0, Alk=2000.0, output=abi.CC_PH_FREE) == 7.502532746463654
    assert oracle.carbon_chemistry(2000.0, 10.0, 35.0, pH=7.5) == 1315.7136384737507


def test_teos10_density_anchor(oracle):
    # SURVEY App. C acceptance value for SeawaterPolynomials' TEOS-10 at (10 °C, 35, 1 bar)
    assert oracle.lib().orc_teos10_polynomial_approximation(10.0, 35.0, 1.0) == 1026.7797599416178


def test_equilibrium_constants_dickson_2007(oracle):
    # test_gasexchange_carbon_chem.jl:113-123 — (value, expected, atol)
    L = oracle.lib()
    S, Tk = 35.0, 298.15
    Is = L.orc_ionic_strength(S)
    KS = L.orc_KS(Tk, S, Is, 0, 0.0)
    checks = [
        (math.log(L.orc_K0(Tk, S)), -3.5617, 1e-4),
        (math.log10(L.orc_K1(Tk, S, 0, 0.0)), -5.8472, 1e-4),
        (math.log10(L.orc_K2(Tk, S, 0, 0.0)), -8.9660, 1e-4),
        (math.log(L.orc_KB(Tk, S, 0, 0.0)), -19.7964, 1e-4),
        (math.log(L.orc_KW(Tk, S, 0, 0.0)), -30.434, 1e-3),
        (math.log(KS), -2.30, 1e-2),
        (math.log(L.orc_KF(Tk, S, Is, KS, 0, 0.0)), -6.09, 1e-2),
        (math.log(L.orc_KP1(Tk, S, 0, 0.0)), -3.71, 1e-2),
        (math.log(L.orc_KP2(Tk, S, 0, 0.0)), -13.727, 1e-3),
        (math.log(L.orc_KP3(Tk, S, 0, 0.0)), -20.24, 1e-2),
        (math.log(L.orc_KSi(Tk, S, Is)), -21.61, 1e-2),
    ]
    for got, want, atol in checks:
        assert abs(got - want) <= atol, (got, want)


def test_pressure_corrections_zeebe_wolf_gladrow(oracle):
    # :134-143 and :152-153 — order K1 K2 KB KW KS KF KP1 KP2 KP3 KspC KspA at Tk = 298.15, P = 300 bar
    want = [1.30804, 1.21341, 1.38024, 1.23784, 1.21844, 1.13151, 1.14852, 1.27298, 1.32217, 1.52962, 1.47866]
    for which, w in enumerate(want):
        assert abs(oracle.lib().orc_pressure_correction(which, 298.15, 300.0) - w) <= 1e-5


def test_solubility_products(oracle):
    # :145-150
    L = oracle.lib()
    assert abs(math.log10(L.orc_KSP_aragonite(298.15, 35.0, 0, 0.0)) - (-6.1883)) <= 1e-4
    assert abs(math.log10(L.orc_KSP_calcite(298.15, 35.0, 0, 0.0)) - (-6.3693)) <= 1e-4


def test_virial_coefficients(oracle):
    # :170-171
    assert abs(oracle.lib().orc_first_virial(298.15) - (-123.2e-6)) <= 1e-8
    assert abs(oracle.lib().orc_cross_virial(298.15) - 22.5e-6) <= 1e-7


def test_pco2_dickson(oracle):
    # :173-181 — pCO₂(DIC = 2136.242890518708, Alk = 2500, T = 25, S = 35) = 350 ± 0.1
    v = oracle.carbon_chemistry(2136.242890518708, 25.0, 35.0, Alk=2500.0, output=abi.CC_PCO2)
    assert abs(v - 350) <= 0.1


def test_reference_solver_is_bimodal(oracle):
    """with atol = 1e-20 the reference's damped Newton exits only on an exactly-zero
    residual: most physical inputs stop after 3–10 iterations, the rest run all 100."""
    rng = np.random.default_rng(1)
    iters = []
    for _ in range(400):
        T, S = rng.uniform(-2, 35), rng.uniform(20, 40)
        DIC = rng.uniform(1800, 2400)
        Alk = rng.uniform(max(DIC * 1.02, 2000), 2600)
        _, ni, nf = oracle.carbon_chemistry(DIC, T, S, Alk=Alk, return_iters=True)
        iters.append(ni)
    iters = np.array(iters)
    assert ((iters <= 12) | (iters == 100)).all()
    assert 0.02 < (iters == 100).mean() < 0.35


def test_ph_outputs_are_ordered(oracle):
    # free > total > seawater scale concentrations ⇒ pHᶠ > pHᵗ > pHˢ
    f = oracle.carbon_chemistry(2100.0, 15.0, 35.0, Alk=2350.0, output=abi.CC_PH_FREE)
    t = oracle.carbon_chemistry(2100.0, 15.0, 35.0, Alk=2350.0, output=abi.CC_PH_TOTAL)
    s = oracle.carbon_chemistry(2100.0, 15.0, 35.0, Alk=2350.0, output=abi.CC_PH_SEAWATER)
    assert f > t > s and 7.5 < s < 8.5


def test_omega_calcite_magnitude(oracle):
    # surface ocean is super-saturated (Ω ≈ 3–6), deep cold water under pressure much less
    surf = oracle.carbon_chemistry(2050.0, 20.0, 35.0, Alk=2350.0, P=0.5, output=abi.CC_OMEGA_CALCITE)
    deep = oracle.carbon_chemistry(2300.0, 2.0, 34.7, Alk=2400.0, P=400.0, silicate=120.0, output=abi.CC_OMEGA_CALCITE)
    assert 2.5 < surf < 7 and 0.3 < deep < 1.5
