"""Biologically active particles — host-side mirror of src/Particles/ (`BiogeochemicalParticles`, `set!`, the
`update_tendencies!` hook and `step_lagrangian_particles!`) with the sugar-kelp individual model of
src/Models/Individuals/SugarKelp/ (`SugarKelp`, `SugarKelpParticles`, `LinearOptimalTemperatureRange`).

Two launches per stage (csrc/kelp.cu) replace the reference's 14: the scatter of all eight uptake / release terms into
Gⁿ of each particle's nearest cell (`NearestPoint`, tracer_interpolation.jl:16-72) and the forward-Euler step of
(A, N, C).  A user-defined particle biogeochemistry is a Julia callable in the reference; only the model it ships
(SugarKelp) exists behind the C ABI.  `advection`: `None` (the reference's `advection = nothing`: particles stay put);
Lagrangian advection by the resolved flow is Oceananigans' `_advect_particles!` and stays there.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib
from .grids import RectilinearGrid, current_stream_ptr

day = 86400.0


@dataclass
class LinearOptimalTemperatureRange:  # equations.jl:206-214
    lower_optimal: float = 10.0
    upper_optimal: float = 15.0
    lower_gradient: Optional[float] = None  # 1 / (lower_optimal + 1.8)
    upper_gradient: float = -0.25

    def __post_init__(self):
        if self.lower_gradient is None:
            self.lower_gradient = 1 / (self.lower_optimal + 1.8)


@dataclass
class SugarKelp:
    """`SugarKelp(FT; …)` — SugarKelp.jl:36-150; derived defaults are evaluated like the reference's keyword defaults."""
    temperature_limit: LinearOptimalTemperatureRange = field(default_factory=LinearOptimalTemperatureRange)
    growth_rate_adjustment: float = 4.5
    photosynthetic_efficiency: float = 4.15e-5 * 24 * 10 ** 6 / (24 * 60 * 60)
    minimum_carbon_reserve: float = 0.01
    structural_carbon: float = 0.2
    exudation: float = 0.5
    erosion_exponent: float = 0.22
    base_erosion_rate: float = 10 ** -6
    saturation_irradiance: float = 90 * day / (10 ** 6)
    structural_dry_weight_per_area: float = 0.5
    structural_dry_to_wet_weight: float = 0.0785
    carbon_reserve_per_carbon: float = 2.1213
    nitrogen_reserve_per_nitrogen: float = 2.72
    minimum_nitrogen_reserve: float = 0.0126
    maximum_nitrogen_reserve: float = 0.0216
    growth_adjustment_2: Optional[float] = None
    growth_adjustment_1: Optional[float] = None
    maximum_specific_growth_rate: float = 0.18
    structural_nitrogen: float = 0.0146
    photosynthesis_at_ref_temp_1: float = 1.22e-3 * 24
    photosynthesis_at_ref_temp_2: float = 1.3e-3 * 24
    photosynthesis_ref_temp_1: float = 285.0
    photosynthesis_ref_temp_2: float = 288.0
    photoperiod_1: float = 0.85
    photoperiod_2: float = 0.3
    respiration_at_ref_temp_1: float = 2.785e-4 * 24
    respiration_at_ref_temp_2: float = 5.429e-4 * 24
    respiration_ref_temp_1: float = 285.0
    respiration_ref_temp_2: float = 290.0
    photosynthesis_arrhenius_temp: Optional[float] = None
    photosynthesis_low_temp: float = 271.0
    photosynthesis_high_temp: float = 296.0
    photosynthesis_high_arrhenius_temp: float = 1414.87
    photosynthesis_low_arrhenius_temp: float = 4547.89
    respiration_arrhenius_temp: Optional[float] = None
    current_speed_for_0p65_uptake: float = 0.03
    nitrate_half_saturation: float = 4.0
    ammonia_half_saturation: float = 1.3
    maximum_nitrate_uptake: Optional[float] = None
    maximum_ammonia_uptake: Optional[float] = None
    current_1: float = 0.72
    current_2: float = 0.28
    current_3: float = 0.045
    base_activity_respiration_rate: float = 1.11e-4 * 24
    base_basal_respiration_rate: float = 5.57e-5 * 24
    exudation_redfield_ratio: float = math.inf
    adapted_latitude: float = 57.5
    newton_iterations: int = 100  # cap of the β solve (the reference's solver: 1000 with an unreachable tolerance)

    def __post_init__(self):
        s = self
        if s.growth_adjustment_2 is None:
            s.growth_adjustment_2 = 0.039 / (2 * (1 - s.minimum_nitrogen_reserve / s.maximum_nitrogen_reserve))
        if s.growth_adjustment_1 is None:
            s.growth_adjustment_1 = 0.18 / (2 * (1 - s.minimum_nitrogen_reserve / s.maximum_nitrogen_reserve)) - s.growth_adjustment_2
        if s.photosynthesis_arrhenius_temp is None:
            s.photosynthesis_arrhenius_temp = ((1 / s.photosynthesis_ref_temp_1 - 1 / s.photosynthesis_ref_temp_2) ** -1
                                               * math.log(s.photosynthesis_at_ref_temp_2 / s.photosynthesis_at_ref_temp_1))
        if s.respiration_arrhenius_temp is None:
            s.respiration_arrhenius_temp = ((1 / s.respiration_ref_temp_1 - 1 / s.respiration_ref_temp_2) ** -1
                                            * math.log(s.respiration_at_ref_temp_2 / s.respiration_at_ref_temp_1))
        if s.maximum_nitrate_uptake is None:
            s.maximum_nitrate_uptake = 10 / s.structural_dry_weight_per_area * 24 * 14 / (10 ** 6)
        if s.maximum_ammonia_uptake is None:
            s.maximum_ammonia_uptake = 12 / s.structural_dry_weight_per_area * 24 * 14 / (10 ** 6)

    # SugarKelp.jl:162-166
    def required_particle_fields(self):
        return ("A", "N", "C")

    def required_tracers(self):
        return ("u", "v", "w", "T", "NO₃", "NH₄", "PAR")

    def coupled_tracers(self):
        return ("NO₃", "NH₄", "DIC", "O₂", "DOC", "DON", "bPOC", "bPON")

    def c_params(self) -> _lib.obm_sugar_kelp_params:
        p = _lib.obm_sugar_kelp_params()
        tl = self.temperature_limit
        for n in _lib.KELP_DOUBLES:
            setattr(p, n, float(getattr(tl, n) if hasattr(tl, n) else getattr(self, n)))
        p.newton_iterations = int(self.newton_iterations)
        return p

    def summary(self):
        return "SugarKelp biogeochemistry (Broch & Slagstad 2012)"


class BiogeochemicalParticles:
    """`BiogeochemicalParticles(number; grid, biogeochemistry, advection, timestepper = ForwardEuler,
    field_interpolation = NearestPoint(), scalefactors = ones(number))` — Particles.jl:58-115.
    `coupled_tracers`: optional {kelp name: model tracer name} override (SugarKelp.jl:165-167, e.g. DON → "DOM")."""

    def __init__(self, number: int, grid: RectilinearGrid, biogeochemistry=None, advection=None, scalefactors=None,
                 coupled_tracers: Optional[dict] = None):
        if advection is not None:
            raise NotImplementedError("LagrangianAdvection is Oceananigans' `_advect_particles!`; pass advection=None")
        self.biogeochemistry = biogeochemistry if biogeochemistry is not None else SugarKelp()
        if not isinstance(self.biogeochemistry, SugarKelp):
            raise NotImplementedError("only the SugarKelp particle biogeochemistry exists behind the C ABI "
                                      "(a user-defined one is a Julia callable in the reference)")
        self.grid, self.number, self.advection = grid, int(number), advection
        dev = grid.device
        z = lambda: torch.zeros(self.number, dtype=torch.float64, device=dev)  # noqa: E731
        self.x, self.y, self.z = z(), z(), z()
        self.fields = {n: z() for n in self.biogeochemistry.required_particle_fields()}
        self.tendencies = {n: z() for n in self.fields}  # ForwardEuler.tendencies, time_stepping.jl:8-16
        self.scalefactors = (torch.ones(self.number, dtype=torch.float64, device=dev) if scalefactors is None
                             else torch.as_tensor(scalefactors, dtype=torch.float64).to(dev).contiguous())
        if self.scalefactors.numel() != self.number:
            raise ValueError("scalefactors must have one entry per particle")
        self.coupled = dict(coupled_tracers) if coupled_tracers else None

    def __len__(self):
        return self.number

    def set(self, **kwargs):
        """`set!(particles; x, y, z, scalefactors, A, N, C)` — set.jl:5-24 (scalars broadcast)."""
        for n, v in kwargs.items():
            target = {"x": self.x, "y": self.y, "z": self.z, "scalefactors": self.scalefactors}.get(n)
            if target is None:
                target = self.fields[n]
            target.copy_(torch.as_tensor(v, dtype=torch.float64).to(target.device).expand_as(target))
        return self

    # ---- C views ------------------------------------------------------------------------------------------
    def c_particles(self) -> _lib.obm_particles:
        g, q = self.grid, _lib.obm_particles()
        q.n = self.number
        q.x, q.y, q.z = self.x.data_ptr(), self.y.data_ptr(), self.z.data_ptr()
        q.A, q.N, q.C = (self.fields[n].data_ptr() for n in ("A", "N", "C"))
        q.scalefactors = self.scalefactors.data_ptr()
        x0 = g.x[0] if g.x is not None else 0.0
        y0 = g.y[0] if g.y is not None else 0.0
        q.x0, q.dx, q.y0, q.dy = x0 + g.dx / 2, g.dx, y0 + g.dy / 2, g.dy
        code = {"Periodic": _lib.OBM_TOPO_PERIODIC, "Bounded": _lib.OBM_TOPO_BOUNDED, "Flat": _lib.OBM_TOPO_FLAT}
        for c in range(3):
            q.topology[c] = code[g.topology[c]]
        return q

    def c_tracers(self, model) -> _lib.obm_kelp_tracers:
        f = _lib.obm_kelp_tracers()
        vel = getattr(model, "velocities", None) or {}
        for n in ("u", "v", "w"):
            setattr(f, n, vel[n].ptr if n in vel else None)
        aux = model.biogeochemistry.biogeochemical_auxiliary_fields()
        f.T, f.NO3, f.NH4 = model.tracers["T"].ptr, model.tracers["NO₃"].ptr, model.tracers["NH₄"].ptr
        f.PAR = aux["PAR"].ptr
        return f

    def _targets(self, model):
        names = self.biogeochemistry.coupled_tracers()
        out = []
        for n in names:
            target = self.coupled.get(n) if self.coupled is not None else n
            out.append(model.Gn[target].ptr if (target is not None and target in model.Gn) else None)
        return out

    # ---- hooks ----------------------------------------------------------------------------------------------
    def update_tendencies(self, bgc, model, stream: Optional[int] = None):
        """`update_tendencies!(bgc, particles, model)` — update_tracer_tendencies.jl:1-18, all coupled tracers in one launch."""
        cg, p, q, f = self.grid.c_grid(), self.biogeochemistry.c_params(), self.c_particles(), self.c_tracers(model)
        s = stream if stream is not None else current_stream_ptr(self.grid.device)
        rc = _lib.load().obm_kelp_update_tendencies(C.byref(cg), C.byref(p), C.byref(q), C.byref(f),
                                                    _lib.pointer_table(self._targets(model)), float(model.clock.time), s)
        _lib.check(rc, "obm_kelp_update_tendencies")

    def step(self, model, dt: float, stream: Optional[int] = None):
        """`update_lagrangian_particle_properties!` (Particles.jl:150-154): advection (none) then
        `time_step_particle_fields!(::ForwardEuler, …)` — tendencies of A, N, C and the Euler update in one launch."""
        cg, p, q, f = self.grid.c_grid(), self.biogeochemistry.c_params(), self.c_particles(), self.c_tracers(model)
        s = stream if stream is not None else current_stream_ptr(self.grid.device)
        out = _lib.pointer_table([self.tendencies[n].data_ptr() for n in ("A", "N", "C")])
        rc = _lib.load().obm_kelp_step(C.byref(cg), C.byref(p), C.byref(q), C.byref(f), float(model.clock.time), float(dt), out, s)
        _lib.check(rc, "obm_kelp_step")

    def summary(self):
        return f"{self.number} BiogeochemicalParticles with {self.biogeochemistry.summary()}"


def SugarKelpParticles(n: int, grid: RectilinearGrid, kelp_parameters: Optional[dict] = None, **kwargs) -> BiogeochemicalParticles:
    """`SugarKelpParticles(n; grid, kelp_parameters = NamedTuple(), kwargs...)` — SugarKelp.jl:152-160."""
    return BiogeochemicalParticles(n, grid, biogeochemistry=SugarKelp(**(kelp_parameters or {})), **kwargs)
