"""The `Biogeochemistry` wrapper and the plugin hooks Oceananigans calls — host-side mirror of
src/OceanBioME.jl:64-169 — plus `BiogeochemicalModel`, a minimal stand-in for the Oceananigans model
(tracers, clock, Gⁿ, RK3/Euler tracer update) that drives the hooks in the reference's order so the
parity tests read like the reference's own (test/test_NutrientsPlanktonDetritus.jl, test_PISCES.jl).

Julia's `f!(…)` names are spelled `f(…)` here; argument order follows the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import ctypes as C
import math

import torch


from . import _lib
from .grids import CenterField, Field, RectilinearGrid, ZFaceField, current_stream_ptr, fill_z_halos
from .negative_tracers import ScaleNegativeTracers, apply_scalers


class Biogeochemistry:
    """`Biogeochemistry(underlying; light_attenuation, sediment, particles, modifiers)` — OceanBioME.jl:100-120."""

    def __init__(self, underlying_biogeochemistry, light_attenuation=None, sediment=None, particles=None,
                 modifiers=None):
        self.underlying_biogeochemistry = underlying_biogeochemistry
        self.light_attenuation = light_attenuation
        self.sediment = sediment
        self.particles = particles
        self.modifiers = modifiers
        self.fuse_state_update = True  # False: one launch per reference step, in the reference's order
        # callable(label) invoked after each launch group of a stage ("modifiers", "light", "state", "sediment",
        # "tendencies", "sediment_tendencies", "particles") — bench.py records a CUDA event there to time every kernel
        # inside the timed region; None costs nothing
        self.stage_marker = None

    def _mark(self, label):
        if self.stage_marker is not None:
            self.stage_marker(label)

    # ---- forwarding (OceanBioME.jl:122-131) -------------------------------------------------------
    def required_biogeochemical_tracers(self):
        return self.underlying_biogeochemistry.required_biogeochemical_tracers()

    def required_biogeochemical_auxiliary_fields(self):
        return self.underlying_biogeochemistry.required_biogeochemical_auxiliary_fields()

    def biogeochemical_drift_velocity(self, name):
        return self.underlying_biogeochemistry.biogeochemical_drift_velocity(name)

    def biogeochemical_auxiliary_fields(self):
        aux = dict(self.underlying_biogeochemistry.biogeochemical_auxiliary_fields())
        if self.light_attenuation is not None:
            aux.update(self.light_attenuation.biogeochemical_auxiliary_fields())
        return aux

    def chlorophyll(self, model):
        return self.underlying_biogeochemistry.chlorophyll(model)

    def conserved_tracers(self, *args, **kwargs):
        return self.underlying_biogeochemistry.conserved_tracers(*args, **kwargs)

    # ---- update_biogeochemical_state!(bgc, model) — OceanBioME.jl:161-167, fixed order ----------
    def update_biogeochemical_state(self, model, stream: Optional[int] = None):
        # PISCES: its pointwise calcite-saturation solve rides in the launch of the last negative scaling (nothing
        # between the two writes DIC, Alk, Si, T or S — see obm_scale_negative_tracers_calcite_saturation)
        u = self.underlying_biogeochemistry
        epilogue = u.calcite_saturation_arguments(model) if (self.fuse_state_update and hasattr(u, "calcite_saturation_arguments")) else None
        fused = _update_modifiers(model, self.modifiers, stream, epilogue)
        self._mark("modifiers")
        # PISCES: zₑᵤ and the mixed-layer mean PAR ride in the launch of the multi-band PAR scan
        light_done = False
        if self.light_attenuation is not None:
            cs = None
            if (self.fuse_state_update and getattr(self.light_attenuation, "supports_column_state", False)
                    and hasattr(u, "column_light_state")):
                cs = u.column_light_state(model)
            if cs is not None:
                self.light_attenuation.update_biogeochemical_state(model, stream, column_state=cs)
                light_done = True
            else:
                self.light_attenuation.update_biogeochemical_state(model, stream)
        self._mark("light")
        if fused or light_done:
            u.update_biogeochemical_state(model, stream, calcite_saturation_done=fused, light_state_done=light_done)
        else:
            u.update_biogeochemical_state(model, stream)
        self._mark("state")
        if self.sediment is not None:
            self.sediment.update_biogeochemical_state(model, stream)
            self._mark("sediment")

    # ---- update_tendencies!(bgc, model) — OceanBioME.jl:148-152 -----------------------------------
    def update_tendencies(self, model, stream: Optional[int] = None):
        """The seam for the fused tendency kernel: every BGC tendency of every tracer is added to
        `model.timestepper.Gⁿ` in one launch (the per-point callable is reduced to zero(grid));
        then the sediment adds its bottom fluxes (Sediments/tracer_coupling.jl:3-28)."""
        self.underlying_biogeochemistry.compute_tendencies(model.grid, model.tracers,
                                                           self.biogeochemical_auxiliary_fields(), model.Gn,
                                                           accumulate=True, stream=stream, time=model.clock.time)
        self._mark("tendencies")
        if self.sediment is not None:
            self.sediment.update_tendencies(self, model, stream)
            self._mark("sediment_tendencies")
        if self.particles is not None:  # OceanBioME.jl:150: update_tendencies!(bgc, bgc.particles, model)
            self.particles.update_tendencies(self, model, stream)
            self._mark("particles")

    def __call__(self, *args):
        """Per-point callable `bgc(i, j, k, grid, Val(name), clock, fields)` — reduced to zero(grid)
        because `update_tendencies` already added every tendency."""
        return 0.0

    def summary(self):
        return f"Biogeochemical model based on {self.underlying_biogeochemistry.summary()}"

    def __repr__(self):
        s = lambda x: "Nothing" if x is None else (tuple(m.summary() for m in x) if isinstance(x, tuple) else x.summary())  # noqa: E731
        return (f"{self.underlying_biogeochemistry.summary()} \n Light attenuation: {s(self.light_attenuation)}\n"
                f" Sediment: {s(self.sediment)}\n Particles: {s(self.particles)}\n Modifiers: {s(self.modifiers)}")


def _update_modifiers(model, modifiers, stream, calcite=None) -> bool:
    """update_biogeochemical_state!(model, modifiers) with tuple broadcast (OceanBioME.jl:169);
    consecutive ScaleNegativeTracers are fused into one launch, order preserved.  `calcite` (PISCES) is handed to the
    LAST launch when that is a negative scaling; returns whether it was consumed."""
    if modifiers is None:
        return False
    mods = modifiers if isinstance(modifiers, tuple) else (modifiers,)
    run = []
    for m in mods:
        if isinstance(m, ScaleNegativeTracers):
            run.append(m)
            continue
        if run:
            apply_scalers(model, run, stream)
            run = []
        m.update_biogeochemical_state(model, stream)
    if run:
        return apply_scalers(model, run, stream, calcite)
    return False


@dataclass
class Clock:
    time: float = 0.0
    iteration: int = 0
    last_stage_dt: float = float("inf")  # `clock.last_stage_Δt` (read by the sediment hook)
    rk3_gamma: float = 1.0
    rk3_zeta: float = float("nan")


class BiogeochemicalModel:
    """Stand-in for `NonhydrostaticModel(; grid, biogeochemistry)` restricted to what the BGC path
    touches: it materialises `required_biogeochemical_tracers` (+ T, S if asked), owns Gⁿ, and steps
    tracers with the BGC source terms only (no advection / diffusion: that is Oceananigans' job)."""

    RK3 = ((8 / 15, 0.0), (5 / 12, -17 / 60), (3 / 4, -5 / 12))  # (γⁿ, ζⁿ) of Oceananigans' RK3

    ADVECTION = {"UpwindBiased1": _lib.ADV_UPWIND1, "Centered2": _lib.ADV_CENTERED2, "UpwindBiased3": _lib.ADV_UPWIND3,
                 "WENO5": _lib.ADV_WENO5}  # WENO(order = 5), what test/test_sediments.jl:37-80 builds its model with

    def __init__(self, grid: RectilinearGrid, biogeochemistry, extra_tracers=(), timestepper="RungeKutta3",
                 boundary_conditions=None, sinking_advection: Optional[str] = None):
        self.grid = grid
        # advection scheme applied to the biogeochemical drift velocities (Oceananigans' `advection` keyword restricted
        # to what a model without resolved flow needs); None = the host model advects the sinking tracers itself
        if sinking_advection is not None and sinking_advection not in self.ADVECTION:
            raise ValueError(f"sinking_advection must be one of {tuple(self.ADVECTION)} or None")
        self.sinking_advection = sinking_advection
        self._drift = {}
        # {tracer: top flux boundary condition} — e.g. DIC=CarbonDioxideGasExchangeBoundaryCondition()
        self.boundary_conditions = dict(boundary_conditions or {})
        self.biogeochemistry = biogeochemistry
        self.clock = Clock()
        names = list(biogeochemistry.required_biogeochemical_tracers())
        names += [t for t in extra_tracers if t not in names]
        self.tracers = {n: CenterField(grid, n) for n in names}
        # Gⁿ of every tracer in one allocation: clearing them is one memset per stage (one graph node instead of one per tracer)
        self._Gn_slab = torch.zeros((len(names),) + tuple(grid.parent_shape), dtype=torch.float64, device=grid.device)
        self.Gn = {n: Field(grid, self._Gn_slab[i], "G" + n) for i, n in enumerate(names)}
        self.Gm = None  # G⁻, allocated on first RK3 step
        self.timestepper = timestepper

    @property
    def auxiliary_fields(self):
        return self.biogeochemistry.biogeochemical_auxiliary_fields()

    def set(self, **values):
        for n, v in values.items():
            self.tracers[n].set(v)
        return self

    def sinking_tracer_names(self):
        """Tracers whose z-halos a hook reads: those advected by a drift velocity here, and the sediment's sinking fluxes."""
        names = []
        if self.sinking_advection is not None:
            names += [n for n in self.tracers if n not in ("T", "S") and self.drift_velocity_field(n) is not None]
        sed = getattr(self.biogeochemistry, "sediment", None)
        if sed is not None:
            names += [n for n in sed.biogeochemistry.sinking_fluxes() if n not in names]
        return names

    def update_state(self):
        # Oceananigans' update_state! fills the halo regions before the biogeochemistry hooks run; of those halos the
        # path reads only the z-planes below / above the sinking tracers' columns (face value of the open bottom face)
        for n in self.sinking_tracer_names():
            fill_z_halos(self.tracers[n])
        self.biogeochemistry.update_biogeochemical_state(self)

    def compute_tendencies(self):
        self._Gn_slab.zero_()
        self.biogeochemistry.update_tendencies(self)
        if self.sinking_advection is not None:
            self.add_sinking_tendencies()
        for name, bc in self.boundary_conditions.items():
            bc.apply_top(self, name)  # G[i, j, Nz] -= flux / Δz, as Oceananigans' apply_z_bcs! does

    def drift_velocity_field(self, name) -> Optional[Field]:
        """`biogeochemical_drift_velocity(bgc, Val(name)).w` as a z-face field, or None for a tracer that does not sink
        (constant speeds are materialised once, like `setup_velocity_fields`, sinking_velocity_fields.jl:10-35:
        every face but the closed top one)."""
        if name not in self._drift:
            w = self.biogeochemistry.biogeochemical_drift_velocity(name)
            if w is None or isinstance(w, Field):
                self._drift[name] = w
            else:
                f = ZFaceField(self.grid, "w" + name)
                u = getattr(self.biogeochemistry, "underlying_biogeochemistry", self.biogeochemistry)
                is_open = getattr(u, "drift_velocity_open_bottom", lambda n: True)(name)
                for k in range(self.grid.Nz):  # faces 1 … Nz; closed bottom: w (1 − e^{(1−k)/2}), k 1-based (:15-17)
                    f.face_interior[k] = float(w) * (1.0 if is_open else (1 - math.exp((1 - (k + 1)) / 2)))
                self._drift[name] = f
        return self._drift[name]

    def add_sinking_tendencies(self, stream: Optional[int] = None):
        """Gⁿ[c] += −∂z(w c) for every sinking tracer, one launch (csrc/sinking.cu)."""
        names = [n for n in self.tracers if n not in ("T", "S") and self.drift_velocity_field(n) is not None]
        cg = self.grid.c_grid()
        s = stream if stream is not None else current_stream_ptr(self.grid.device)
        for c0 in range(0, len(names), _lib.OBM_MAX_SINKING_TRACERS):
            chunk = names[c0:c0 + _lib.OBM_MAX_SINKING_TRACERS]
            rc = _lib.load().obm_sinking_tendencies(
                C.byref(cg), len(chunk), _lib.pointer_table([self.tracers[n].ptr for n in chunk]),
                _lib.pointer_table([self.drift_velocity_field(n).ptr for n in chunk]),
                _lib.pointer_table([self.Gn[n].ptr for n in chunk]), self.ADVECTION[self.sinking_advection], 1, s)
            _lib.check(rc, "obm_sinking_tendencies")

    # ---- run!(simulation) for column ensembles (many independent 1-D models stepped on the device) ------------
    def run(self, dt: float, steps: int, graph: bool = False, output_every: int = 0, output_names=None):
        """Integrate `steps` time steps; returns {name: tensor (n_outputs, Nz, Ny, Nx)} of interior snapshots taken every
        `output_every` steps (device-resident).

        `graph=True`: the FIRST step runs eagerly (it alone sees the pre-run `last_stage_dt`), then ONE time step — every
        stage's negative scaling, PAR scan, fused tendencies, sinking, sediment hooks, boundary fluxes and tracer update —
        is captured in a CUDA graph and replayed for the rest of the run: ≈ 15 launches per step with no host work between
        them, whatever the number of columns.  Needs kernel arguments that do not depend on the clock: NPZD / LOBSTER
        family (not PISCES: its day lengths are host-evaluated per stage), a constant or field surface PAR, no particles.
        Bit-identical to the eager loop."""
        names = list(output_names or [n for n in self.tracers if n not in ("T", "S")])
        nout = steps // output_every if output_every else 0
        shape = (nout, self.grid.Nz, self.grid.Ny, self.grid.Nx)
        out = {n: torch.empty(shape, dtype=torch.float64, device=self.grid.device) for n in names}

        def snapshot(it):
            if output_every and (it + 1) % output_every == 0:
                for n in names:
                    out[n][(it + 1) // output_every - 1].copy_(self.tracers[n].interior)

        if not graph or steps < 3:
            for it in range(steps):
                self.time_step(dt)
                snapshot(it)
            return out
        bgc = self.biogeochemistry
        u = getattr(bgc, "underlying_biogeochemistry", bgc)
        if getattr(u, "clock_dependent_parameters", False):
            raise ValueError("graph=True needs kernel parameters that do not depend on the clock (not PISCES)")
        if getattr(bgc, "particles", None) is not None:
            raise ValueError("graph=True: particles are stepped with host logic between the launches; use the eager loop")
        if callable(getattr(getattr(bgc, "light_attenuation", None), "surface_PAR", None)):
            raise ValueError("graph=True: a surface PAR function of time is evaluated on the host; use a constant or a field")
        self.time_step(dt)
        snapshot(0)
        dev = self.grid.device
        sediment = getattr(bgc, "sediment", None)
        state = [self.tracers, self.Gn, self.Gm, self.auxiliary_fields] + ([sediment.fields] if sediment is not None else [])
        for extra in ("Gn", "Gm", "tendencies", "previous_tendencies"):
            d = getattr(sediment, extra, None) if sediment is not None else None
            if isinstance(d, dict):
                state.append(d)
        saved = [(f, f.data.clone()) for d in state if d is not None for f in d.values() if hasattr(f, "data")]
        clock = (self.clock.time, self.clock.iteration, self.clock.last_stage_dt)
        sed_host = (sediment.last_dt, sediment.iteration) if sediment is not None and hasattr(sediment, "last_dt") else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up: lazily built fields, allocator
            self.time_step(dt)
        torch.cuda.current_stream(dev).wait_stream(side)
        for f, v in saved:
            f.data.copy_(v)
        self.clock.time, self.clock.iteration, self.clock.last_stage_dt = clock
        if sed_host is not None:  # the sediment's own stepper keeps Δt of its last call and a counter on the host
            sediment.last_dt, sediment.iteration = sed_host
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.time_step(dt)
        if sed_host is not None:
            sediment.iteration = sed_host[1] + (sediment.iteration - sed_host[1]) * (steps - 1)
        # the capture does not execute; the host-side clock advanced once while capturing: rewind, then count the replays
        self.clock.time, self.clock.iteration, self.clock.last_stage_dt = clock
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(1, steps):
            g.replay()
            snapshot(it)
        e1.record()
        self._graph, self.replay_events = g, (e0, e1)
        self.clock.time = clock[0] + (steps - 1) * dt
        self.clock.iteration = clock[1] + steps - 1
        stages = self.RK3 if self.timestepper != "Euler" else ((1.0, 0.0),)
        self.clock.last_stage_dt = dt * (stages[-1][0] + stages[-1][1])
        return out

    def _substep(self, dt, gamma, zeta):
        """U += Δt(γGⁿ + ζG⁻), G⁻ ← Gⁿ for every tracer in one launch (csrc/timestepping.cu)."""
        names = list(self.tracers)
        tab = lambda d: _lib.pointer_table([d[n].ptr for n in names])  # noqa: E731
        cg = self.grid.c_grid()
        rc = _lib.load().obm_rk3_substep(C.byref(cg), len(names), tab(self.tracers), tab(self.Gn),
                                         tab(self.Gm) if self.Gm is not None else None, float(dt), float(gamma),
                                         float(zeta or 0.0), int(bool(zeta)), int(self.Gm is not None),
                                         current_stream_ptr(self.grid.device))
        _lib.check(rc, "obm_rk3_substep")

    def step_lagrangian_particles(self):
        """`step_lagrangian_particles!(model, Δt)` with Δt the stage just completed.  Oceananigans calls it at the END of
        every stage, after `update_state!` has recomputed the tendencies; this stand-in calls `update_state` at the START
        of a stage, so the same interleaving (substep → state → tendencies → particles → substep …) is obtained by
        stepping the particles right after the tendencies — the last stage's particle step of a run is then pending
        until the next stage, or `finish_particles()`."""
        p = getattr(self.biogeochemistry, "particles", None)
        if p is not None and self.clock.last_stage_dt != float("inf"):
            p.step(self, self.clock.last_stage_dt)

    def finish_particles(self):
        """Bring the particles to the model time at the end of a run (state, tendencies, pending particle step)."""
        self.update_state()
        self.compute_tendencies()
        self.step_lagrangian_particles()
        self.clock.last_stage_dt = float("inf")

    def time_step(self, dt: float):
        if self.timestepper == "Euler":
            self.update_state()
            self.compute_tendencies()
            self.step_lagrangian_particles()
            self._substep(dt, 1.0, None)
            self.clock.time += dt
            self.clock.last_stage_dt = dt
        else:
            if self.Gm is None:
                self.Gm = {n: CenterField(self.grid, "G⁻" + n) for n in self.tracers}
            for stage, (gamma, zeta) in enumerate(self.RK3):
                self.clock.rk3_gamma, self.clock.rk3_zeta = gamma, (zeta if zeta else float("nan"))
                self.update_state()
                self.compute_tendencies()
                self.step_lagrangian_particles()
                self._substep(dt, gamma, zeta)
                self.clock.time += dt * (gamma + zeta)
                self.clock.last_stage_dt = dt * (gamma + zeta)
        self.clock.iteration += 1
