"""GPU parity: the x–y gas-exchange kernel (through the C ABI) against the oracle on seeded surface
states, every operand form of the reference (number / (x, y, t) function / 2-D field / 3-D field), the
reference's own known answers (test/test_gasexchange_carbon_chem.jl:31,58-69), the top-cell tendency
update, sub-range launches and error behaviour.  Tolerance: 1e-10 relative on the CO₂ flux (it inherits
the pCO₂ of the carbonate solve), 1e-12 on the O₂ flux, both scale-aware against k·max(|water|, |air|)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib as abi
from oceanbiome_b200 import gas_exchange as ge
from helpers import RTOL_CARBON, RTOL_TENDENCY

pytestmark = pytest.mark.gpu


def host(f):
    return np.ascontiguousarray(f.data.cpu().numpy())


def surface_state(cuda, size=(70, 9, 5), seed=0):
    grid = ob.RectilinearGrid(size=size, extent=(700.0, 90.0, 50.0), device=cuda)
    rng = np.random.default_rng(seed)
    shp = grid.parent_shape
    vals = {"T": rng.uniform(-1.5, 32, shp), "S": rng.uniform(28, 39, shp), "DIC": rng.uniform(1850, 2350, shp),
            "O₂": rng.uniform(120, 380, shp), "Si": rng.uniform(0, 120, shp), "PO₄": rng.uniform(0, 3, shp)}
    vals["Alk"] = vals["DIC"] * rng.uniform(1.04, 1.18, shp)
    fields = {}
    for n, v in vals.items():
        fields[n] = ob.CenterField(grid, n)
        fields[n].data.copy_(torch.from_numpy(v))
    return grid, fields, vals, rng


def scale_err(got, want, scale):
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(scale), 1e-300)))


class _Clock:
    time = 0.0


def test_reference_known_answers(cuda):
    # test_gasexchange_carbon_chem.jl:19-31: size (1, 1, 2) grid, T = 15, S = 35, DIC = 2220, Alk = 2500
    grid = ob.RectilinearGrid(size=(1, 1, 2), extent=(1, 1, 1), device=cuda)
    for air in (413.1, (lambda x, y, t: 413.0), ob.CenterField(grid, "air", fill=413.0)):
        bc = ob.CarbonDioxideGasExchangeBoundaryCondition(air_concentration=air)
        bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem())
        model = ob.BiogeochemicalModel(grid, bgc, extra_tracers=("T", "S"), boundary_conditions={"DIC": bc})
        model.set(T=15.0, S=35.0, DIC=2220.0, Alk=2500.0)
        value = bc.getbc(model).interior[0, 0, 0].item()
        assert isinstance(bc.condition.func, ob.GasExchange)
        assert abs(value - (-8e-6)) <= 1e-6
        model.time_step(1.0)
        assert torch.isfinite(model.tracers["DIC"].interior).all()


def test_two_carbonate_systems(cuda):
    # :35-69 — two DIC/Alk pairs, each with its own boundary condition
    grid = ob.RectilinearGrid(size=(1, 1, 2), extent=(1, 1, 1), device=cuda)
    bc1 = ob.CarbonDioxideGasExchangeBoundaryCondition(water_concentration=ob.CarbonDioxideConcentration(DIC="DIC1", Alk="Alk1"))
    bc2 = ob.CarbonDioxideGasExchangeBoundaryCondition(water_concentration=ob.CarbonDioxideConcentration(DIC="DIC2", Alk="Alk2"))
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(2))
    model = ob.BiogeochemicalModel(grid, bgc, extra_tracers=("T", "S"), boundary_conditions={"DIC1": bc1, "DIC2": bc2})
    model.set(T=15.0, S=35.0, DIC1=2220.0, Alk1=2500.0, DIC2=2221.0, Alk2=2501.0)
    v1, v2 = (bc.getbc(model).interior[0, 0, 0].item() for bc in (bc1, bc2))
    assert abs(v1 + 8e-6) <= 1e-6 and abs(v2 + 8e-6) <= 1e-6 and v1 != v2
    model.time_step(1.0)


@pytest.mark.parametrize("k660", ["Ho06", "Wanninkhof09", "McGillis01", "Nightingale00"])
def test_co2_flux_matches_oracle(cuda, oracle, k660):
    grid, fields, vals, rng = surface_state(cuda, seed=1)
    og = oracle.Grid.like(grid)
    wind = ob.Field2D(grid, "u10")
    air = ob.Field2D(grid, "pCO2air")
    hw, ha = rng.uniform(0, 18, og.plane_shape), rng.uniform(370, 460, og.plane_shape)
    wind.data.copy_(torch.from_numpy(hw))
    air.data.copy_(torch.from_numpy(ha))
    tv = ob.SchmidtScaledTransferVelocity(schmidt_number=ob.CarbonDioxidePolynomialSchmidtNumber(),
                                          base_transfer_velocity=getattr(ge, k660)(),
                                          solubility=ge.MolPerKgPerAtmToMMolPerCubicMPerMicroAtm())
    g = ob.CarbonDioxideGasExchangeBoundaryCondition(transfer_velocity=tv, wind_speed=wind,
                                                     air_concentration=air).condition.func
    got = host(g.compute_flux(grid, _Clock, fields))
    p = g.c_params()
    want = oracle.gas_exchange_flux(og, p, vals["T"], vals["S"], DIC=vals["DIC"], Alk=vals["Alk"], wind_speed=hw,
                                    air_concentration=ha)
    gi, wi = og.interior(got)[0], og.interior(want)[0]
    # scale: k·max(pCO₂, air) — the difference pCO₂ − air cancels
    kk = np.vectorize(lambda u, T, S: oracle.transfer_velocity(p, u, T, S))(
        og.interior(hw)[0], og.interior(vals["T"])[-1], og.interior(vals["S"])[-1])
    scale = kk * np.maximum(og.interior(ha)[0], np.abs(wi / np.where(kk > 0, kk, 1) + og.interior(ha)[0]))
    assert scale_err(gi, wi, np.maximum(scale, np.abs(wi))) <= RTOL_CARBON
    assert not (got - og_zero_interior(og, got)).any()  # halos of the flux plane untouched


def og_zero_interior(og, a):
    b = a.copy()
    og.interior(b)[...] = 0
    return a - b


def test_silicate_phosphate_operands(cuda, oracle):
    grid, fields, vals, _ = surface_state(cuda, seed=2)
    og = oracle.Grid.like(grid)
    # names → fields
    g = ob.CarbonDioxideGasExchangeBoundaryCondition(silicate_and_phosphate_names=("Si", "PO₄"), wind_speed=6.5,
                                                     air_concentration=400).condition.func
    got = og.interior(host(g.compute_flux(grid, _Clock, fields)))[0]
    p = g.c_params()
    p.wind_speed, p.air_concentration = 6.5, 400.0
    want = og.interior(oracle.gas_exchange_flux(og, p, vals["T"], vals["S"], DIC=vals["DIC"], Alk=vals["Alk"],
                                                silicate=vals["Si"], phosphate=vals["PO₄"]))[0]
    assert np.max(np.abs(got - want)) <= RTOL_CARBON * np.max(np.abs(want)) * 50
    # NamedTuple of values
    g2 = ob.CarbonDioxideGasExchangeBoundaryCondition(silicate_and_phosphate_names={"silicate": 40.0, "phosphate": 1.5},
                                                      wind_speed=6.5, air_concentration=400).condition.func
    got2 = og.interior(host(g2.compute_flux(grid, _Clock, fields)))[0]
    p2 = g2.c_params()
    p2.wind_speed, p2.air_concentration = 6.5, 400.0
    want2 = og.interior(oracle.gas_exchange_flux(og, p2, vals["T"], vals["S"], DIC=vals["DIC"], Alk=vals["Alk"]))[0]
    assert np.max(np.abs(got2 - want2)) <= RTOL_CARBON * np.max(np.abs(want2)) * 50
    assert not np.array_equal(got, got2)


def test_o2_flux_matches_oracle_and_top_tendency(cuda, oracle):
    grid, fields, vals, rng = surface_state(cuda, size=(130, 7, 4), seed=3)
    og = oracle.Grid.like(grid)
    g = ob.OxygenGasExchangeBoundaryCondition(wind_speed=lambda x, y, t: 3.0 + 0.01 * x + 0.02 * y + t).condition.func

    class clock:
        time = 0.5

    G = ob.CenterField(grid, "GO2")
    hG = rng.uniform(-1e-5, 1e-5, og.parent_shape)
    G.data.copy_(torch.from_numpy(hG))
    flux = ob.Field2D(grid, "flux")
    g.compute_flux(grid, clock, fields, flux=flux, G_top=G)
    p = g.c_params()
    p.air_concentration = 9352.7
    xs = (np.arange(grid.Nx) + 0.5) * grid.dx
    ys = (np.arange(grid.Ny) + 0.5) * grid.dy
    hw = np.zeros(og.plane_shape)
    og.interior(hw)[0][...] = 3.0 + 0.01 * xs[None, :] + 0.02 * ys[:, None] + 0.5
    hG2 = hG.copy()
    want = oracle.gas_exchange_flux(og, p, vals["T"], vals["S"], tracer=vals["O₂"], wind_speed=hw, G_top=hG2)
    gi, wi = og.interior(host(flux))[0], og.interior(want)[0]
    scale = np.maximum(np.abs(wi), 1e-7 * 400)
    assert scale_err(gi, wi, scale) <= RTOL_TENDENCY * 100  # water − α·air cancels up to ~1/100
    gG = host(G)
    assert np.array_equal(og.interior(gG)[:-1], og.interior(hG)[:-1])  # only the top cell is touched
    assert np.max(np.abs(og.interior(gG)[-1] - og.interior(hG2)[-1])) <= 1e-12 * np.max(np.abs(og.interior(hG2)[-1]))
    halo = gG.copy()
    og.interior(halo)[...] = 0
    ref = hG.copy()
    og.interior(ref)[...] = 0
    assert np.array_equal(halo, ref)


def test_3d_air_field_and_subrange(cuda, oracle):
    grid, fields, vals, rng = surface_state(cuda, seed=4)
    og = oracle.Grid.like(grid)
    air3 = ob.CenterField(grid, "air3")
    ha3 = rng.uniform(380, 440, og.parent_shape)
    air3.data.copy_(torch.from_numpy(ha3))
    g = ob.CarbonDioxideGasExchangeBoundaryCondition(air_concentration=air3, wind_speed=9.0).condition.func
    p = g.c_params()
    p.wind_speed = 9.0
    ha = np.ascontiguousarray(ha3[grid.Nz - 1 + grid.Hz][None])
    want = oracle.gas_exchange_flux(og, p, vals["T"], vals["S"], DIC=vals["DIC"], Alk=vals["Alk"], air_concentration=ha)
    full = host(g.compute_flux(grid, _Clock, fields))
    assert np.max(np.abs(full - want)) <= RTOL_CARBON * np.max(np.abs(want)) * 50
    # two partial launches over j ranges give the same plane (slab pipelining / multi-GPU slabs)
    flux = ob.Field2D(grid, "f")
    with grid.restrict(0, 4):
        g.compute_flux(grid, _Clock, fields, flux=flux)
    part = host(flux).copy()
    assert np.array_equal(og.interior(part)[0][:4], og.interior(full)[0][:4]) and not og.interior(part)[0][4:].any()
    with grid.restrict(4, grid.Ny):
        g.compute_flux(grid, _Clock, fields, flux=flux)
    assert np.array_equal(host(flux), full)


def test_nan_and_extreme_inputs_propagate(cuda, oracle):
    grid, fields, vals, _ = surface_state(cuda, size=(8, 2, 3), seed=5)
    og = oracle.Grid.like(grid)
    top = grid.Nz - 1 + grid.Hz
    vals["T"][top, grid.Hy, grid.Hx] = np.nan
    vals["O₂"][top, grid.Hy, grid.Hx + 1] = np.inf
    vals["O₂"][top, grid.Hy, grid.Hx + 2] = 0.0
    for n in ("T", "O₂"):
        fields[n].data.copy_(torch.from_numpy(vals[n]))
    g = ob.OxygenGasExchangeBoundaryCondition(wind_speed=0.0).condition.func
    got = og.interior(host(g.compute_flux(grid, _Clock, fields)))[0]
    p = g.c_params()
    p.air_concentration, p.wind_speed = 9352.7, 0.0
    want = og.interior(oracle.gas_exchange_flux(og, p, vals["T"], vals["S"], tracer=vals["O₂"]))[0]
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got[0, 0]) and np.isnan(got[0, 1])  # 0·Inf
    m = ~np.isnan(want)
    assert np.array_equal(got[m], want[m])  # u₁₀ = 0 ⇒ exactly ±0 everywhere else


def test_error_behaviour(cuda):
    lib = abi.load()
    grid, fields, _, _ = surface_state(cuda, size=(4, 2, 2))
    cg = grid.c_grid()
    p = ob.OxygenGasExchangeBoundaryCondition().condition.func.c_params()
    flux = ob.Field2D(grid, "f")
    args = lambda **kw: [kw.get(k) for k in ("T", "S", "tracer", "DIC", "Alk", "sil", "phos", "wind", "air", "flux", "G")]  # noqa: E731
    T, S, O2 = fields["T"].ptr, fields["S"].ptr, fields["O₂"].ptr
    assert lib.obm_gas_exchange_flux(C.byref(cg), None, *args(T=T, S=S, tracer=O2, flux=flux.ptr), None) == -1
    assert lib.obm_gas_exchange_flux(C.byref(cg), C.byref(p), *args(T=T, S=S, flux=flux.ptr), None) == -1  # no tracer
    assert b"tracer" in lib.obm_last_error()
    assert lib.obm_gas_exchange_flux(C.byref(cg), C.byref(p), *args(T=T, S=S, tracer=O2), None) == -1  # no output
    p.water_kind = abi.OBM_GE_WATER_PCO2
    assert lib.obm_gas_exchange_flux(C.byref(cg), C.byref(p), *args(T=T, S=S, flux=flux.ptr), None) == -1  # no DIC/Alk
    p.water_kind = 7
    assert lib.obm_gas_exchange_flux(C.byref(cg), C.byref(p), *args(T=T, S=S, tracer=O2, flux=flux.ptr), None) == -3
    p.water_kind, p.k660_order = abi.OBM_GE_WATER_TRACER, 4
    assert lib.obm_gas_exchange_flux(C.byref(cg), C.byref(p), *args(T=T, S=S, tracer=O2, flux=flux.ptr), None) == -2
    p.k660_order = 2
    assert lib.obm_gas_exchange_flux(C.byref(cg), C.byref(p), *args(T=T, S=S, tracer=O2, flux=flux.ptr), None) == 0
    torch.cuda.synchronize()
