"""Air–sea gas exchange — host-side mirror of src/Models/GasExchange/ (SURVEY §8 row f-1).

Same constructors and argument meaning as the reference (GasExchange.jl:68-147):
`GasExchangeBoundaryCondition`, `CarbonDioxideGasExchangeBoundaryCondition`,
`OxygenGasExchangeBoundaryCondition`, `SchmidtScaledTransferVelocity`, the k₆₆₀ family and the two
Schmidt-number polynomials.  The returned `GasExchange` object is what the reference stores in
`FluxBoundaryCondition(...).condition.func`; here it evaluates *all* surface columns in one launch
of csrc/gas_exchange.cu (`compute_flux`) and can apply the flux to the top-cell tendency (`apply`).
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .carbon_chemistry import CarbonChemistry
from .grids import Field, Field2D, RectilinearGrid, current_stream_ptr

hour = 3600.0  # Oceananigans.Units.hour


class PolynomialParameterisation:
    """`PolynomialParameterisation{N}` — generic_parameterisations.jl:1-36."""

    def __init__(self, order: int, coefficients: Sequence[float]):
        if len(coefficients) != order + 1:
            raise ValueError("You must provide N+1 coefficients for an order N polynomial")
        self.order = int(order)
        self.coefficients = tuple(float(c) for c in coefficients)

    def summary(self):
        return f"Order {self.order} polynomial parameterisation"

    def __repr__(self):
        return f"{self.summary()}\n    p(x) = Σ{{n ∈ Z : [0, {self.order}]}}(cₙ xⁿ⁻¹) where c = {self.coefficients}"


# Schmidt numbers — schmidt_number.jl:6-16 (Wanninkhof 2014)
def CarbonDioxidePolynomialSchmidtNumber(a=2116.8, b=-136.25, c=4.7353, d=-0.092307, e=0.0007555):
    return PolynomialParameterisation(4, (a, b, c, d, e))


def OxygenPolynomialSchmidtNumber(a=1920.4, b=-135.6, c=5.2122, d=-0.10939, e=0.00093777):
    return PolynomialParameterisation(4, (a, b, c, d, e))


# k₆₆₀ family — gas_transfer_velocity.jl:47-135
def Wanninkhof99(scale_factor=0.0283 / hour / 100):
    return PolynomialParameterisation(3, (0, 0, 0, scale_factor))


def Ho06(scale_factor=0.266 / hour / 100):
    return PolynomialParameterisation(2, (0, 0, scale_factor))


def Nightingale00(linear=0.333 / hour / 100, quadratic=0.222 / hour / 100):
    return PolynomialParameterisation(2, (0, linear, quadratic))


def McGillis01(constant=3.3 / hour / 100, cubic=0.026 / hour / 100):
    return PolynomialParameterisation(3, (constant, 0, 0, cubic))


def Sweeny07(scale_factor=0.27 / hour / 100):
    return PolynomialParameterisation(2, (0, 0, scale_factor))


def Wanninkhof09(constant=3 / hour / 100, linear=0.1 / hour / 100, quadratic=0.064 / hour / 100,
                 cubic=0.011 / hour / 100):
    return PolynomialParameterisation(3, (constant, linear, quadratic, cubic))


def Wanninkhof14(scale_factor=0.251 / hour / 100):
    return PolynomialParameterisation(2, (0, 0, scale_factor))


def ERA5(scale_factor=0.270875 / hour / 100):
    return PolynomialParameterisation(2, (0, 0, scale_factor))


def JRA55(scale_factor=0.2601975 / hour / 100):
    return PolynomialParameterisation(2, (0, 0, scale_factor))


def NCEP1(scale_factor=0.2866424 / hour / 100):
    return PolynomialParameterisation(2, (0, 0, scale_factor))


def CCMP2(scale_factor=0.256789 / hour / 100):
    return PolynomialParameterisation(2, (0, 0, scale_factor))


class MolPerKgPerAtmToMMolPerCubicMPerMicroAtm:
    """K0(T + 273.15, S)·ρ(T, S)/10³ of the carbon chemistry model — gas_solubility.jl:52-65."""

    def __init__(self, carbon_chemistry: Optional[CarbonChemistry] = None):
        self.carbon_chemistry = carbon_chemistry or CarbonChemistry()


class SchmidtScaledTransferVelocity:
    """k(u₁₀, T, S) = k₆₆₀(u₁₀) / √(Sc(T)/660) · solubility(T, S) — gas_transfer_velocity.jl:11-33."""

    def __init__(self, *, schmidt_number: PolynomialParameterisation,
                 base_transfer_velocity: Optional[PolynomialParameterisation] = None, solubility=None):
        self.base_transfer_velocity = base_transfer_velocity or Ho06()
        self.schmidt_number = schmidt_number
        self.solubility = solubility  # None ≙ (T, S) -> 1
        if self.base_transfer_velocity.order > 3:
            raise ValueError("base_transfer_velocity: polynomial order ≤ 3 supported")
        if self.schmidt_number.order != 4:
            raise ValueError("schmidt_number must be an order-4 PolynomialParameterisation")
        if solubility is not None and not isinstance(solubility, MolPerKgPerAtmToMMolPerCubicMPerMicroAtm):
            raise TypeError("solubility must be None or MolPerKgPerAtmToMMolPerCubicMPerMicroAtm")

    def summary(self):
        return "SchmidtScaledTransferVelocity{PolynomialParameterisation, PolynomialParameterisation}"


class Wanninkhof92Solubility:
    """Ostwald solubility β(T, S)/Tk — gas_solubility.jl:27-47."""

    def __init__(self, A1, A2, A3, B1, B2, B3):
        self.coefficients = tuple(float(v) for v in (A1, A2, A3, B1, B2, B3))


def OxygenSolubility(A1=-58.3877, A2=85.8079, A3=23.8439, B1=-0.034892, B2=0.015578, B3=-0.0019387):
    return Wanninkhof92Solubility(A1, A2, A3, B1, B2, B3)


class PartiallySolubleGas:
    """`PartiallySolubleGas(; air_concentration, solubility)` — gas_solubility.jl:1-25."""

    def __init__(self, *, air_concentration, solubility: Wanninkhof92Solubility):
        self.air_concentration = air_concentration
        self.solubility = solubility


class OxygenConcentration:
    """Model tracer `O₂` at the surface — GasExchange.jl:32-37."""
    tracer = "O₂"

    def summary(self):
        return "Model tracer `OxygenConcentration`"


class TracerConcentration(OxygenConcentration):
    """The `[Tracer]Concentration` a user builds for another tracer (GasExchange.jl:62-64)."""

    def __init__(self, tracer: str):
        self.tracer = tracer


class CarbonDioxideConcentration:
    """pCO₂ of the surface cell from the carbon chemistry model — carbon_dioxide_concentration.jl:15-60."""

    def __init__(self, *, carbon_chemistry: Optional[CarbonChemistry] = None, air_pressure: float = 1.0,
                 silicate_and_phosphate_names=None, DIC: str = "DIC", Alk: str = "Alk"):
        self.carbon_chemistry = carbon_chemistry or CarbonChemistry()
        self.air_pressure = air_pressure  # stored, never read by the reference's surface_value either
        self.silicate_and_phosphate_names = silicate_and_phosphate_names
        self.DIC, self.Alk = DIC, Alk

    def summary(self):
        return f"`CarbonChemistry` derived partial pressure of CO₂ (pCO₂) {{{self.DIC}, {self.Alk}}}"


def _surface_operand(value, grid: RectilinearGrid, clock, what: str):
    """`surface_value` (surface_values.jl:1-39): number | (x, y, t) function | Field → (plane ptr|None, scalar)."""
    if isinstance(value, Field):
        if value.is_2d:
            return value.ptr, 0.0, value
        # 3-D field: its surface level is a contiguous plane of the parent
        kt = grid.Nz - 1 + grid.Hz
        plane = value.data[kt]
        return plane.data_ptr(), 0.0, plane
    if callable(value):
        f = Field2D(grid, what)
        x0 = grid.x[0] if grid.x is not None else 0.0
        y0 = grid.y[0] if grid.y is not None else 0.0
        x = (x0 + (torch.arange(grid.Nx, dtype=torch.float64, device=grid.device) + 0.5) * grid.dx).view(1, -1)
        y = (y0 + (torch.arange(grid.Ny, dtype=torch.float64, device=grid.device) + 0.5) * grid.dy).view(-1, 1)
        v = value(x, y, clock.time if clock is not None else 0.0)
        f.interior.copy_(torch.as_tensor(v, dtype=torch.float64, device=grid.device).expand(grid.Ny, grid.Nx))
        return f.ptr, 0.0, f
    return None, float(value), None


class GasExchange:
    """`GasExchange(wind_speed, transfer_velocity, water_concentration, air_concentration)` —
    gas_exchange.jl:19-38, evaluated for all (i, j) at once."""

    def __init__(self, wind_speed, transfer_velocity: SchmidtScaledTransferVelocity, water_concentration,
                 air_concentration):
        self.wind_speed = wind_speed
        self.transfer_velocity = transfer_velocity
        self.water_concentration = water_concentration
        self.air_concentration = air_concentration

    def summary(self):
        return f"Air-sea `GasExchange` model for {type(self.water_concentration).__name__}"

    # -- C parameters ------------------------------------------------------------------------------------------
    def c_params(self) -> _lib.obm_gas_exchange_params:
        p = _lib.obm_gas_exchange_params()
        tv = self.transfer_velocity
        p.k660_order = tv.base_transfer_velocity.order
        for n, c in enumerate(tv.base_transfer_velocity.coefficients):
            p.k660[n] = c
        for n, c in enumerate(tv.schmidt_number.coefficients):
            p.schmidt[n] = c
        p.solubility_kind = _lib.OBM_GE_SOLUBILITY_ONE if tv.solubility is None else _lib.OBM_GE_SOLUBILITY_K0_RHO
        wc = self.water_concentration
        if isinstance(wc, CarbonDioxideConcentration):
            p.water_kind = _lib.OBM_GE_WATER_PCO2
            p.carbon_chemistry = wc.carbon_chemistry.c_params()
            sp = wc.silicate_and_phosphate_names
            if sp is not None:
                p.use_silicate_phosphate = 1
                if isinstance(sp, dict):  # NamedTuple of values
                    p.silicate, p.phosphate = (float(v) for v in sp.values())
        else:
            p.water_kind = _lib.OBM_GE_WATER_TRACER
        ac = self.air_concentration
        if isinstance(ac, PartiallySolubleGas):
            p.air_kind = _lib.OBM_GE_AIR_WANNINKHOF92
            for n, c in enumerate(ac.solubility.coefficients):
                p.w92[n] = c
        else:
            p.air_kind = _lib.OBM_GE_AIR_PLAIN
        return p

    # -- evaluation --------------------------------------------------------------------------------------------
    def compute_flux(self, grid: RectilinearGrid, clock, model_fields: dict, flux: Optional[Field] = None,
                     G_top: Optional[Field] = None, stream: Optional[int] = None) -> Field:
        """flux[i, j] = g(i, j, grid, clock, model_fields) for every surface column; optionally
        G_top[i, j, Nz] -= flux / Δz (what Oceananigans does with a top flux boundary condition)."""
        p = self.c_params()
        keep = []
        wptr, p.wind_speed, k1 = _surface_operand(self.wind_speed, grid, clock, "wind_speed")
        ac = self.air_concentration
        inner = ac.air_concentration if isinstance(ac, PartiallySolubleGas) else ac
        aptr, p.air_concentration, k2 = _surface_operand(inner, grid, clock, "air_concentration")
        keep += [k1, k2]
        wc = self.water_concentration
        f = lambda n: model_fields[n].ptr  # noqa: E731
        tracer = DIC = Alk = sil = phos = None
        if isinstance(wc, CarbonDioxideConcentration):
            DIC, Alk = f(wc.DIC), f(wc.Alk)
            sp = wc.silicate_and_phosphate_names
            if sp is not None and not isinstance(sp, dict):
                sil, phos = f(sp[0]), f(sp[1])
        else:
            tracer = f(wc.tracer)
        if flux is None and G_top is None:
            flux = Field2D(grid, "gas_exchange_flux")
        cg = grid.c_grid()
        s = stream if stream is not None else current_stream_ptr(grid.device)
        rc = _lib.load().obm_gas_exchange_flux(C.byref(cg), C.byref(p), f("T"), f("S"), tracer, DIC, Alk, sil, phos,
                                               wptr, aptr, flux.ptr if flux is not None else None,
                                               G_top.ptr if G_top is not None else None, s)
        _lib.check(rc, "obm_gas_exchange_flux")
        del keep
        return flux

    def __call__(self, grid, clock, model_fields, **kw):
        return self.compute_flux(grid, clock, model_fields, **kw)


class FluxBoundaryCondition:
    """Minimal stand-in for Oceananigans' `FluxBoundaryCondition(func; discrete_form = true)`:
    `.condition.func` is the `GasExchange`, as the reference's tests expect
    (test_gasexchange_carbon_chem.jl:30)."""

    class _Condition:
        def __init__(self, func):
            self.func = func

    def __init__(self, func: GasExchange):
        self.condition = FluxBoundaryCondition._Condition(func)

    def apply_top(self, model, tracer_name: str, stream: Optional[int] = None):
        fields = dict(model.tracers)
        self.condition.func.compute_flux(model.grid, model.clock, fields, G_top=model.Gn[tracer_name], stream=stream)

    def getbc(self, model) -> Field:
        return self.condition.func.compute_flux(model.grid, model.clock, dict(model.tracers))


def GasExchangeBoundaryCondition(*, water_concentration, air_concentration, transfer_velocity, wind_speed):
    """GasExchange.jl:68-82."""
    return FluxBoundaryCondition(GasExchange(wind_speed, transfer_velocity, water_concentration, air_concentration))


def CarbonDioxideGasExchangeBoundaryCondition(*, carbon_chemistry: Optional[CarbonChemistry] = None,
                                              transfer_velocity: Optional[SchmidtScaledTransferVelocity] = None,
                                              air_concentration=413, wind_speed=2, water_concentration=None,
                                              silicate_and_phosphate_names=None):
    """GasExchange.jl:103-123 (defaults: Ho06 k₆₆₀, Wanninkhof-2014 CO₂ Schmidt number, K0·ρ solubility)."""
    carbon_chemistry = carbon_chemistry or CarbonChemistry()
    if transfer_velocity is None:
        transfer_velocity = SchmidtScaledTransferVelocity(
            schmidt_number=CarbonDioxidePolynomialSchmidtNumber(),
            solubility=MolPerKgPerAtmToMMolPerCubicMPerMicroAtm(carbon_chemistry))
    if water_concentration is None:
        water_concentration = CarbonDioxideConcentration(carbon_chemistry=carbon_chemistry,
                                                         silicate_and_phosphate_names=silicate_and_phosphate_names)
    return GasExchangeBoundaryCondition(water_concentration=water_concentration, air_concentration=air_concentration,
                                        transfer_velocity=transfer_velocity, wind_speed=wind_speed)


def OxygenGasExchangeBoundaryCondition(*, transfer_velocity: Optional[SchmidtScaledTransferVelocity] = None,
                                       water_concentration=None, air_concentration=None, wind_speed=2):
    """GasExchange.jl:138-145."""
    if transfer_velocity is None:
        transfer_velocity = SchmidtScaledTransferVelocity(schmidt_number=OxygenPolynomialSchmidtNumber())
    if water_concentration is None:
        water_concentration = OxygenConcentration()
    if air_concentration is None:
        air_concentration = PartiallySolubleGas(air_concentration=9352.7, solubility=OxygenSolubility())
    return GasExchangeBoundaryCondition(water_concentration=water_concentration, air_concentration=air_concentration,
                                        transfer_velocity=transfer_velocity, wind_speed=wind_speed)
