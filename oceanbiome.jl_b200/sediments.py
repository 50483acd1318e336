"""Bottom sediments — host-side mirror of src/Sediments/ + src/Models/Sediments/: `InstantRemineralisationSediment`,
`SimpleMultiGSediment`, `BiogeochemicalSediment` with the two hooks Oceananigans calls
(`update_biogeochemical_state!(model, sediment)`, `update_tendencies!(bgc, sediment, model)`), each ONE fused launch
(csrc/sediments.cu) instead of the reference's ≈ 15 tiny :xy kernels.

Parity note: the reference leaves this path untested (test/test_sediments.jl:106-163 is commented out) and depends on
Oceananigans' flux operator / time-stepper internals, so the kernel is checked against the oracle under the stated
assumptions (first-order upwind or centred face value; step → cache → recompute) and on total-nitrogen conservation.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from .grids import Field, Field2D, RectilinearGrid, ZFaceField, current_stream_ptr, require_cuda

day = 86400.0


@dataclass
class InstantRemineralisation:  # instant_remineralisation.jl:13-19,83-96
    burial_efficiency_constant1: float = 0.013
    burial_efficiency_constant2: float = 0.53
    burial_efficiency_half_saturation: float = 7.0 / 6.56
    sinking_tracers: Sequence[str] = ("P", "D")
    remineralisation_reciever: str = "N"

    def required_sediment_fields(self):
        return ("storage",)

    def required_tracers(self):
        return ()

    def sinking_fluxes(self):
        return tuple(self.sinking_tracers)

    def coupled_tracers(self):
        return (self.remineralisation_reciever,)

    def summary(self):
        return "Single-layer instant remineralisation (Float64)"


@dataclass
class SimpleMultiG:  # simple_multi_G.jl:15-38,104-132
    sinking_redfield: Optional[float] = 6.56
    fast_decay_rate: float = 2 / day
    slow_decay_rate: float = 0.2 / day
    fast_redfield: float = 0.1509
    slow_redfield: float = 0.13
    fast_fraction: float = 0.74
    slow_fraction: float = 0.26
    refactory_fraction: float = 0.1
    sedimentation_rate: float = 1.0
    anoxia_half_saturation: float = 1.0
    nitrate_oxidation_params: tuple = (-1.9785, 0.2261, -0.0615, -0.0289, -0.36109, -0.0232)
    denitrification_params: tuple = (-3.0790, 1.7509, 0.0593, -0.1923, 0.0604, 0.0662)
    anoxic_params: tuple = (-3.9476, 2.6269, -0.2426, -1.3349, 0.1826, -0.0143)
    solid_dep_params: tuple = (0.233, 0.336, 982.0, -1.548)
    sinking_nitrogen: Sequence[str] = ("sPOM", "bPOM")
    sinking_carbon: Optional[Sequence[str]] = None

    @property
    def carbon(self):
        return self.sinking_redfield is None

    def required_sediment_fields(self):
        return ("Ns", "Nf", "Nr", "Cs", "Cf", "Cr") if self.carbon else ("Ns", "Nf", "Nr")

    def required_tracers(self):
        return ("NO₃", "NH₄", "O₂")

    def sinking_fluxes(self):
        return tuple(self.sinking_nitrogen) + (tuple(self.sinking_carbon) if self.carbon else ())

    def coupled_tracers(self):
        return ("NO₃", "NH₄", "O₂", "DIC") if self.carbon else ("NO₃", "NH₄", "O₂")

    def summary(self):
        return "Single-layer multi-G sediment model (Float64)"


class BiogeochemicalSediment:
    """`BiogeochemicalSediment(grid, biogeochemistry; timestepper = :QuasiAdamsBashforth2)` — Sediments.jl:14-61."""

    def __init__(self, grid: RectilinearGrid, biogeochemistry, timestepper: str = "QuasiAdamsBashforth2",
                 advection: str = "UpwindBiased1", bottom_height: Optional[Field] = None, chi: float = 0.1):
        self.grid, self.biogeochemistry = grid, biogeochemistry
        if timestepper not in ("QuasiAdamsBashforth2", "RungeKutta3"):
            raise ValueError(f"{timestepper} is not configured for sediment models")  # timesteppers.jl:10
        self.timestepper, self.advection, self.chi = timestepper, advection, chi
        self.fields = {n: Field2D(grid, n) for n in biogeochemistry.required_sediment_fields()}
        self.Gn = {n: Field2D(grid, "Gⁿ" + n) for n in self.fields}
        self.Gm = {n: Field2D(grid, "G⁻" + n) for n in self.fields}
        names = tuple(biogeochemistry.required_tracers()) + tuple(biogeochemistry.sinking_fluxes())
        self.tracked_fields = {n: Field2D(grid, n) for n in names}
        self.bottom_indices = None  # OneField (bottom_indices.jl:5)
        if bottom_height is not None:
            self.bottom_indices = calculate_bottom_indices(grid, bottom_height)
        self.last_dt = math.inf
        self.iteration = 0

    def prognostic_fields(self):
        return self.fields

    # -- C structs -----------------------------------------------------------------------------------------
    def c_params(self) -> _lib.obm_sediment_params:
        b, p = self.biogeochemistry, _lib.obm_sediment_params()
        # the bottom face: every upwind-biased scheme (UpwindBiased 1 / 3, WENO5) is first-order there — its higher-order
        # stencils would leave the interior (sinking.cu::face_value) — so only Centered2 differs
        p.advection = _lib.ADV_CENTERED2 if self.advection == "Centered2" else _lib.ADV_UPWIND1
        p.timestepper = _lib.TS_AB2 if self.timestepper == "QuasiAdamsBashforth2" else _lib.TS_RK3
        if isinstance(b, InstantRemineralisation):
            p.model = _lib.SED_INSTANT_REMINERALISATION
            p.nsinking_nitrogen = len(b.sinking_tracers)
            p.burial_efficiency_constant1 = b.burial_efficiency_constant1
            p.burial_efficiency_constant2 = b.burial_efficiency_constant2
            p.burial_efficiency_half_saturation = b.burial_efficiency_half_saturation
        else:
            p.model = _lib.SED_SIMPLE_MULTI_G
            p.carbon = 1 if b.carbon else 0
            p.nsinking_nitrogen = len(b.sinking_nitrogen)
            p.nsinking_carbon = len(b.sinking_carbon) if b.carbon else 0
            p.sinking_redfield = 0.0 if b.carbon else float(b.sinking_redfield)
            for k in ("fast_decay_rate", "slow_decay_rate", "fast_redfield", "slow_redfield", "fast_fraction",
                      "slow_fraction", "refactory_fraction", "sedimentation_rate", "anoxia_half_saturation"):
                setattr(p, k, float(getattr(b, k)))
            for k, n in (("nitrate_oxidation_params", 6), ("denitrification_params", 6), ("anoxic_params", 6), ("solid_dep_params", 4)):
                for q in range(n):
                    getattr(p, k)[q] = float(getattr(b, k)[q])
        return p

    def c_fields(self, model, bgc) -> _lib.obm_sediment_fields:
        b, f = self.biogeochemistry, _lib.obm_sediment_fields()
        f.bottom_indices_xy = self.bottom_indices.data_ptr() if self.bottom_indices is not None else None
        t = model.tracers
        if b.required_tracers():
            f.NO3, f.NH4, f.O2 = t["NO₃"].ptr, t["NH₄"].ptr, t["O₂"].ptr
        for n, name in enumerate(b.sinking_fluxes()):
            f.sinking[n] = t[name].ptr
            f.sinking_w[n] = self._w_field(bgc, name).ptr
        for n, name in enumerate(self.fields):
            f.pools[n], f.Gn[n], f.Gm[n] = self.fields[name].ptr, self.Gn[name].ptr, self.Gm[name].ptr
        for n, name in enumerate(self.tracked_fields):
            f.tracked_xy[n] = self.tracked_fields[name].ptr
        for n, name in enumerate(b.coupled_tracers()):
            f.G_coupled[n] = model.Gn[name].ptr if name in model.Gn else None
        return f

    def _w_field(self, bgc, name) -> Field:
        """`biogeochemical_drift_velocity(model.biogeochemistry, Val(name)).w` as a z-face field
        (constant speeds are materialised once, like `setup_velocity_fields`)."""
        cache = self.__dict__.setdefault("_w", {})
        if name not in cache:
            w = bgc.biogeochemical_drift_velocity(name)
            if isinstance(w, Field):
                cache[name] = w
            else:
                f = ZFaceField(self.grid, "w" + name)
                u = getattr(bgc, "underlying_biogeochemistry", bgc)
                is_open = getattr(u, "drift_velocity_open_bottom", lambda n: True)(name)
                for k in range(self.grid.Nz):  # sinking_velocity_fields.jl:15-17; a closed bottom feeds the sediment nothing
                    f.face_interior[k] = (0.0 if w is None else float(w)) * (1.0 if is_open else (1 - math.exp((1 - (k + 1)) / 2)))
                cache[name] = f
        return cache[name]

    # -- hooks -------------------------------------------------------------------------------------------------
    def update_biogeochemical_state(self, model, stream: Optional[int] = None):
        """update_state.jl:6-16 with Δt = model.clock.last_stage_Δt (∞ before the first step ⇒ no pool update)."""
        bgc = model.biogeochemistry
        dt = getattr(model.clock, "last_stage_dt", math.inf)
        p, f, cg = self.c_params(), self.c_fields(model, bgc), self.grid.c_grid()
        chi = -0.5 if (dt != self.last_dt) else self.chi  # Oceananigans' AB2 takes an Euler step when Δt changed
        s = stream if stream is not None else current_stream_ptr(self.grid.device)
        rc = _lib.load().obm_sediment_update_state(C.byref(cg), C.byref(p), C.byref(f), float(dt), chi, s)
        _lib.check(rc, "obm_sediment_update_state")
        if math.isfinite(dt):
            self.last_dt = dt
            self.iteration += 1

    def update_tendencies(self, bgc, model, stream: Optional[int] = None):
        """tracer_coupling.jl:3-28"""
        p, f, cg = self.c_params(), self.c_fields(model, bgc), self.grid.c_grid()
        s = stream if stream is not None else current_stream_ptr(self.grid.device)
        rc = _lib.load().obm_sediment_update_tendencies(C.byref(cg), C.byref(p), C.byref(f), s)
        _lib.check(rc, "obm_sediment_update_tendencies")

    def summary(self):
        return f"`BiogeochemicalSediment` with {self.biogeochemistry.summary()}"


def calculate_bottom_indices(grid: RectilinearGrid, bottom_height: Field, stream=None) -> torch.Tensor:
    """`calculate_bottom_indices(grid::ImmersedBoundaryGrid)` (bottom_indices.jl:19-26) for a grid-fitted bottom;
    returns the 2-D Int64 parent array of 1-based bottom-cell indices."""
    require_cuda(bottom_height)
    out = torch.ones(grid.plane_shape, dtype=torch.int64, device=grid.device)
    cg = grid.c_grid()
    s = stream if stream is not None else current_stream_ptr(grid.device)
    rc = _lib.load().obm_find_bottom_cells(C.byref(cg), bottom_height.ptr, out.data_ptr(), s)
    _lib.check(rc, "obm_find_bottom_cells")
    return out


def InstantRemineralisationSediment(grid, sinking_tracers=("P", "D"), remineralisation_reciever="N",
                                    burial_efficiency_constant1=0.013, burial_efficiency_constant2=0.53,
                                    burial_efficiency_half_saturation=7.0 / 6.56, **kwargs):
    """instant_remineralisation.jl:83-96"""
    return BiogeochemicalSediment(grid, InstantRemineralisation(burial_efficiency_constant1, burial_efficiency_constant2,
                                                                burial_efficiency_half_saturation, tuple(sinking_tracers),
                                                                remineralisation_reciever), **kwargs)


def SimpleMultiGSediment(grid, sinking_nitrogen=("sPOM", "bPOM"), sinking_carbon=None, sinking_redfield="default",
                         sedimentation_rate=None, timestepper="QuasiAdamsBashforth2", advection="UpwindBiased1",
                         bottom_height=None, chi=0.1, **params):
    """simple_multi_G.jl:104-132; sedimentation_rate defaults to 982 |z₁|^(−1.548) (cm/year)."""
    if sinking_redfield == "default":
        sinking_redfield = 6.56 if sinking_carbon is None else None
    if sedimentation_rate is None:
        sedimentation_rate = 982 * abs(float(grid.zc[0])) ** (-1.548)
    b = SimpleMultiG(sinking_redfield=sinking_redfield, sedimentation_rate=sedimentation_rate,
                     sinking_nitrogen=tuple(sinking_nitrogen), sinking_carbon=tuple(sinking_carbon) if sinking_carbon else None,
                     **params)
    return BiogeochemicalSediment(grid, b, timestepper=timestepper, advection=advection, bottom_height=bottom_height, chi=chi)
