"""CPU: the gas-exchange oracle against the reference's own known answers
(test/test_gasexchange_carbon_chem.jl:31,164-194) and the host-side mirror of its constructors."""
import math

import numpy as np
import pytest

import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib as abi
from oceanbiome_b200 import gas_exchange as ge


def co2_params(**kw):
    g = ob.CarbonDioxideGasExchangeBoundaryCondition(**kw).condition.func
    p = g.c_params()
    p.air_concentration, p.wind_speed = float(g.air_concentration), float(g.wind_speed)
    return p


def o2_params(**kw):
    g = ob.OxygenGasExchangeBoundaryCondition(**kw).condition.func
    p = g.c_params()
    p.air_concentration, p.wind_speed = float(g.air_concentration.air_concentration), float(g.wind_speed)
    return p


def test_reference_flux_value(oracle):
    # test_gasexchange_carbon_chem.jl:24-31: T = 15, S = 35, DIC = 2220, Alk = 2500, air 413.1 → ≈ −8e-6 ± 1e-6
    p = co2_params(air_concentration=413.1)
    v = oracle.gas_exchange_point(p, 15.0, 35.0, DIC=2220.0, Alk=2500.0)
    assert abs(v - (-8e-6)) <= 1e-6
    # :64-69 the second carbonate system differs
    v2 = oracle.gas_exchange_point(p, 15.0, 35.0, DIC=2221.0, Alk=2501.0)
    assert abs(v2 - (-8e-6)) <= 1e-6 and v2 != v


def test_schmidt_numbers_wanninkhof_2014(oracle):
    # :164-165
    assert abs(oracle.polynomial(ge.CarbonDioxidePolynomialSchmidtNumber().coefficients, 20.0) - 668) <= 1
    assert abs(oracle.polynomial(ge.OxygenPolynomialSchmidtNumber().coefficients, 20.0) - 568) <= 1


def test_virial_coefficients_dickson_2007(oracle):
    # :169-170
    L = oracle.lib()
    assert abs(L.orc_first_virial(25 + 273.15) - (-123.2e-6)) <= 1e-8
    assert abs(L.orc_cross_virial(25 + 273.15) - 22.5e-6) <= 1e-7


def test_pco2_and_po2_values(oracle):
    # :179-185
    pco2 = oracle.carbon_chemistry(2136.242890518708, 25.0, 35.0, 2500.0, output=abi.CC_PCO2)
    assert abs(pco2 - 350) <= 0.1
    p = o2_params()
    po2 = p.air_concentration * oracle.w92_solubility(list(p.w92), 25.0, 35.0)
    assert abs(po2 - 200) <= 50


def test_polynomial_forms(oracle):
    # generic_parameterisations.jl:24-36 — left-to-right sums of cₙ·xⁿ (x², x³ as products, x⁴ compensated)
    c = (1.5, -2.25, 0.3, 0.07, -0.011)
    for x in (0.0, 1.0, -3.7, 12.25, 31.4159):
        for order in range(5):
            want = c[0]
            terms = [c[1] * x, c[2] * (x * x), c[3] * (x * x * x), c[4] * float(np.float64(x) ** 4)]
            for n in range(order):
                want = want + terms[n]
            got = oracle.polynomial(c[:order + 1], x)
            assert got == pytest.approx(want, rel=2e-16, abs=0), (order, x)


def test_k660_family_values():
    # gas_transfer_velocity.jl:52-135 in cm/hr at u₁₀ = 10 m/s: scale·u² (or cubic) · 3600·100
    u = 10.0
    cmhr = 3600.0 * 100

    def val(p):
        return sum(c * u ** n for n, c in enumerate(p.coefficients)) * cmhr

    assert val(ge.Ho06()) == pytest.approx(26.6)
    assert val(ge.Wanninkhof99()) == pytest.approx(28.3)
    assert val(ge.Nightingale00()) == pytest.approx(3.33 + 22.2)
    assert val(ge.McGillis01()) == pytest.approx(3.3 + 26.0)
    assert val(ge.Sweeny07()) == pytest.approx(27.0)
    assert val(ge.Wanninkhof09()) == pytest.approx(3 + 1 + 6.4 + 11)
    assert val(ge.Wanninkhof14()) == pytest.approx(25.1)
    assert val(ge.ERA5()) == pytest.approx(27.0875)
    assert val(ge.JRA55()) == pytest.approx(26.01975)
    assert val(ge.NCEP1()) == pytest.approx(28.66424)
    assert val(ge.CCMP2()) == pytest.approx(25.6789)
    with pytest.raises(ValueError):
        ge.PolynomialParameterisation(2, (1, 2))


def test_transfer_velocity_composition(oracle):
    # k = k660(u)/√(Sc/660)·K0·ρ/1000 (gas_transfer_velocity.jl:32-33, gas_solubility.jl:65)
    p = co2_params()
    T, S, u = 12.0, 34.0, 7.5
    k660 = 0.266 / 3600.0 / 100 * (u * u)
    Sc = oracle.polynomial(ge.CarbonDioxidePolynomialSchmidtNumber().coefficients, T)
    L = oracle.lib()
    sol = L.orc_K0(T + 273.15, S) * L.orc_teos10_polynomial_approximation(T, S, 0.0) / 1000.0
    want = k660 / math.sqrt(Sc / 660.0) * sol
    assert oracle.transfer_velocity(p, u, T, S) == pytest.approx(want, rel=1e-15)
    po = o2_params()
    Sco = oracle.polynomial(ge.OxygenPolynomialSchmidtNumber().coefficients, T)
    assert oracle.transfer_velocity(po, u, T, S) == pytest.approx(k660 / math.sqrt(Sco / 660.0), rel=1e-15)


def test_gridded_flux_and_top_tendency(oracle):
    rng = np.random.default_rng(5)
    zf = np.concatenate([[-13.0, -11.0, -9.0], np.array([-7.0, -4.0, -2.0, -0.5, 0.0]), [0.5, 1.0, 1.5]])
    zc = 0.5 * (zf[:-1] + zf[1:])
    g = oracle.Grid(6, 5, 4, 3, 3, 3, zc, zf)
    shp = g.parent_shape
    T, S = rng.uniform(0, 30, shp), rng.uniform(30, 38, shp)
    DIC = rng.uniform(1900, 2300, shp)
    Alk = DIC * rng.uniform(1.05, 1.15, shp)
    wind = rng.uniform(0, 15, g.plane_shape)
    air = rng.uniform(380, 450, g.plane_shape)
    p = co2_params()
    G = np.zeros(shp)
    flux = oracle.gas_exchange_flux(g, p, T, S, DIC=DIC, Alk=Alk, wind_speed=wind, air_concentration=air, G_top=G)
    fi, Gi = g.interior(flux)[0], g.interior(G)
    # per-point agreement with the scalar entry
    Ti, Si, Di, Ai = (g.interior(a)[-1] for a in (T, S, DIC, Alk))
    wi, ai = g.interior(wind)[0], g.interior(air)[0]
    for (j, i) in ((0, 0), (2, 3), (4, 5)):
        assert fi[j, i] == oracle.gas_exchange_point(p, Ti[j, i], Si[j, i], DIC=Di[j, i], Alk=Ai[j, i], u10=wi[j, i],
                                                     air=ai[j, i])
    # only the top cell's tendency is touched: G[Nz] = −flux/Δz(Nz), Δz = 0.5
    assert np.array_equal(Gi[-1], -(fi / 0.5))
    assert not Gi[:-1].any()
    # halos untouched
    assert flux.sum() == fi.sum()
    # oxygen: tracer value − α·air
    O2 = rng.uniform(150, 350, shp)
    po = o2_params()
    fo = oracle.gas_exchange_flux(g, po, T, S, tracer=O2)
    Oi = g.interior(O2)[-1]
    assert g.interior(fo)[0][1, 2] == oracle.gas_exchange_point(po, Ti[1, 2], Si[1, 2], tracer=Oi[1, 2])
    # sign: supersaturated water loses gas (positive = upward flux)
    assert oracle.gas_exchange_point(po, 25.0, 35.0, tracer=1000.0) > 0 > oracle.gas_exchange_point(po, 25.0, 35.0, tracer=0.0)


def test_silicate_phosphate_forms(oracle):
    # carbon_dioxide_concentration.jl:62-64: nothing → (0, 0); NamedTuple → values; names → fields
    p0 = co2_params()
    assert p0.use_silicate_phosphate == 0
    pv = co2_params(silicate_and_phosphate_names={"silicate": 50.0, "phosphate": 2.0})
    assert pv.use_silicate_phosphate == 1 and (pv.silicate, pv.phosphate) == (50.0, 2.0)
    base = oracle.gas_exchange_point(p0, 15.0, 35.0, DIC=2220.0, Alk=2500.0)
    with_sp = oracle.gas_exchange_point(pv, 15.0, 35.0, DIC=2220.0, Alk=2500.0, silicate=50.0, phosphate=2.0)
    assert with_sp != base and abs(with_sp - base) < 1e-6  # nutrient alkalinity raises pCO₂ a little


def test_constructor_structure():
    bc = ob.CarbonDioxideGasExchangeBoundaryCondition()
    g = bc.condition.func
    assert isinstance(g, ob.GasExchange)  # test_gasexchange_carbon_chem.jl:30
    assert isinstance(g.water_concentration, ob.CarbonDioxideConcentration)
    assert g.air_concentration == 413 and g.wind_speed == 2
    assert "GasExchange" in g.summary()
    o = ob.OxygenGasExchangeBoundaryCondition().condition.func
    assert isinstance(o.water_concentration, ob.OxygenConcentration)
    assert isinstance(o.air_concentration, ob.PartiallySolubleGas) and o.air_concentration.air_concentration == 9352.7
    p = o.c_params()
    assert (p.water_kind, p.air_kind, p.solubility_kind) == (abi.OBM_GE_WATER_TRACER, abi.OBM_GE_AIR_WANNINKHOF92,
                                                            abi.OBM_GE_SOLUBILITY_ONE)
    two = ob.CarbonDioxideConcentration(DIC="DIC2", Alk="Alk2")
    assert (two.DIC, two.Alk) == ("DIC2", "Alk2")


def test_oxygen_flux_against_an_independent_restatement(oracle):
    """The whole O₂ exchange retyped from the reference in plain Python (gas_exchange.jl:26-38,
    gas_transfer_velocity.jl:32-33 with Ho06, schmidt_number.jl:13-16, gas_solubility.jl:34-47 incl. its B2-used-twice
    quadratic term, GasExchange.jl: air 9352.7 mmol O₂ m⁻³, wind 2 m/s): no coefficient is taken from the package."""
    p = o2_params()
    hour = 3600.0
    for T, S, O2, u in ((25.0, 35.0, 200.0, 2.0), (4.0, 33.1, 330.0, 9.5), (15.5, 36.7, 260.0, 0.3)):
        Sc = 1920.4 + -135.6 * T + 5.2122 * T ** 2 + -0.10939 * T ** 3 + 0.00093777 * T ** 4
        k = (0.266 / hour / 100) * u ** 2 / math.sqrt(Sc / 660)
        Tk = T + 273.15
        Tk100 = Tk / 100
        beta = math.exp(-58.3877 + 85.8079 / Tk100 + 23.8439 * math.log(Tk100) + S * (-0.034892 + 0.015578 * Tk100 + 0.015578 * Tk100 ** 2))
        want = k * (O2 - 9352.7 * (beta / Tk))
        got = oracle.gas_exchange_point(p, T, S, tracer=O2, u10=u)
        assert math.isclose(got, want, rel_tol=1e-13), (T, S, got, want)


def test_carbon_dioxide_flux_composition_against_an_independent_restatement(oracle):
    """CO₂: k₆₆₀(u) / √(Sc(T)/660) · K₀(T, S) ρ(T, S) / 10³ · (pCO₂ − air) with the Schmidt polynomial (Wanninkhof 2014)
    and Ho06 retyped here; pCO₂, K₀ and ρ come from the carbonate oracle, which the reference's docstring goldens pin
    bit for bit (tests/test_oracle_carbon.py)."""
    p = co2_params(air_concentration=413.0)
    L = oracle.lib()
    for T, S, DIC, Alk, u in ((15.0, 35.0, 2220.0, 2500.0, 2.0), (3.0, 34.0, 2150.0, 2300.0, 11.0), (27.0, 36.5, 1950.0, 2350.0, 6.0)):
        Sc = 2116.8 + -136.25 * T + 4.7353 * T ** 2 + -0.092307 * T ** 3 + 0.0007555 * T ** 4
        k = (0.266 / 3600.0 / 100) * u ** 2 / math.sqrt(Sc / 660) * (L.orc_K0(T + 273.15, S) * L.orc_teos10_polynomial_approximation(T, S, 0.0) / 10 ** 3)
        pCO2 = oracle.carbon_chemistry(DIC=DIC, T=T, S=S, Alk=Alk, output=abi.CC_PCO2)
        got = oracle.gas_exchange_point(p, T, S, DIC=DIC, Alk=Alk, u10=u)
        assert math.isclose(got, k * (pCO2 - 413.0), rel_tol=1e-12), (T, got, k * (pCO2 - 413.0))
