"""`BoxModel` — host-side mirror of src/BoxModel/boxmodel.jl:35-160 and timesteppers.jl:1-95, re-designed as a
device-resident ENSEMBLE driver.

The reference integrates one 0-D box at a time on the CPU (`BoxModelGrid()` = a `(Flat, Flat, Flat)` grid,
≈ 8 µs per RK stage, benchmark/box_model.jl:67).  Here `BoxModelGrid(n)` is n independent boxes laid along x, and
one time step is a handful of launches whatever n is: the same fused kernels as the 3-D models evaluate all
tendencies of all boxes (`update_biogeochemical_state`, `update_tendencies`), `obm_rk3_substep` updates every
tracer and caches G⁻ in one launch, and — when nothing but the tabulated time series depends on time — the whole
step is captured once in a CUDA graph and replayed (`run(..., graph=True)`), so parameter / initial-condition
sweeps and calibration ensembles run entirely on the device.

Same keyword surface as the reference: `BoxModel(biogeochemistry=…, grid=…, forcing=…, timestepper=…,
clock=…, prescribed_tracers=…)`, `set(model, **values)` (`set!`), `time_step(Δt)` (`time_step!`), `run` (`run!`).
`forcing[name]` and `prescribed_tracers[name]` are functions of time `f(t)` (numbers or per-box tensors).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional

import torch

from . import _lib
from .biogeochemistry import Clock
from .grids import CenterField, Field, RectilinearGrid, current_stream_ptr


class SpeedyOutput:
    """`SpeedyOutput(filename; overwrite_existing = true)` — src/BoxModel/output_writer.jl:1-56: the time series of
    every model field, one entry per output, appended while the model runs and read back with `load_output`.
    The reference writes a JLD2 file (`timeseries/<name>/<iteration>`); this mirror writes a NumPy `.npz` with the same
    content — `t` and one array (n_outputs, n_boxes) per field — because JLD2 is a Julia serialisation format.
    Snapshots are gathered on the device and copied to the host once, when the file is written."""

    def __init__(self, filename: str, overwrite_existing: bool = True):
        import os
        if os.path.exists(filename) and not overwrite_existing:
            raise FileExistsError(filename)
        self.filename, self.overwrite_existing = filename, overwrite_existing
        self._t, self._series = [], {}

    def __call__(self, model):
        """One output: the reference's `(save::SpeedyOutput)(simulation)`."""
        self._t.append(float(model.clock.time))
        for n, f in model.fields.items():
            self._series.setdefault(n, []).append(f.interior.reshape(-1).clone())

    def flush(self):
        import numpy as np
        data = {n: torch.stack(v).cpu().numpy() for n, v in self._series.items()}
        data["t"] = np.asarray(self._t, dtype=np.float64)
        with open(self.filename, "wb") as fh:  # np.savez would append ".npz" to a bare name
            np.savez(fh, **data)


def load_output(save: SpeedyOutput, name=None):
    """`load_output(save)` / `load_output(save, name)` — output_writer.jl:34-56: dict name → array sorted by time."""
    import numpy as np
    with np.load(save.filename) as z:
        order = np.argsort(z["t"])
        if name is not None:
            return z[str(name)][order]
        return {n: z[n][order] for n in z.files}


def BoxModelGrid(n: int = 1, device="cuda", z: Optional[float] = None) -> RectilinearGrid:
    """`BoxModelGrid(; z)` (src/OceanBioME.jl:171) for `n` independent boxes: x is the ensemble axis, y and z are Flat.
    The single level is centred on `z` (the reference's Flat z-node, e.g. `BoxModelGrid(; z = -5)` in test_PISCES.jl:37);
    without it the level spans z ∈ [−1, 0] so that kernels reading z see a finite node."""
    span = (-1.0, 0.0) if z is None else (float(z) - 0.5, float(z) + 0.5)
    return RectilinearGrid(size=(int(n),), x=(0.0, float(n)), z=span, topology=("Periodic", "Flat", "Flat"), halo=(0,),
                           device=device)


class BoxModel:
    """`BoxModel(; biogeochemistry, grid, forcing, timestepper, clock, prescribed_tracers)` — boxmodel.jl:35-90."""

    RK3 = ((8 / 15, None), (5 / 12, -17 / 60), (3 / 4, -5 / 12))  # (γⁿ, ζⁿ), Oceananigans' RungeKutta3TimeStepper

    def __init__(self, biogeochemistry, grid: Optional[RectilinearGrid] = None, forcing: Optional[dict] = None,
                 timestepper: str = "RungeKutta3", clock: Optional[Clock] = None,
                 prescribed_tracers: Optional[Dict[str, Callable]] = None, fused_step: bool = False):
        """`fused_step=True` (NPZD / LOBSTER family without sediment or particles): every stage is ONE launch —
        `compute_tendencies!`, `rk3_substep!` and `cache_previous_tendencies!` (timesteppers.jl:30-93) evaluated per
        cell from one read of its tracers (`obm_npd_tendencies_substep`); the tracers and G⁻ are those of the
        three-launch path bit for bit, Gⁿ is not materialised (`model.Gn` then only carries the forcing)."""
        self.biogeochemistry = biogeochemistry
        self.grid = grid if grid is not None else BoxModelGrid()
        self.clock = clock or Clock()
        self.prescribed_tracers = dict(prescribed_tracers or {})
        names = list(biogeochemistry.required_biogeochemical_tracers())
        aux = biogeochemistry.biogeochemical_auxiliary_fields()
        # a prescribed name is a tracer (e.g. T) or an auxiliary field the biogeochemistry reads (e.g. PAR)
        names += [n for n in self.prescribed_tracers if n not in names and n not in aux]
        self.fields = {n: CenterField(self.grid, n) for n in names}
        self.forcing = {n: (forcing or {}).get(n) for n in names}
        unknown = set(forcing or {}) - set(names)
        if unknown:
            raise ValueError(f"forcing for unknown tracers {sorted(unknown)}")
        if timestepper not in ("RungeKutta3", "Euler"):
            raise ValueError("timestepper must be 'RungeKutta3' or 'Euler'")
        self.timestepper = timestepper
        # prognostic = everything the biogeochemistry steps; prescribed tracers are overwritten every stage
        self.prognostic = [n for n in names if n not in self.prescribed_tracers]
        # Gⁿ of every field in one allocation: a step without forcing clears all of them with one memset
        self._Gn_slab = torch.zeros((len(names),) + tuple(self.grid.parent_shape), dtype=torch.float64,
                                    device=self.grid.device)
        self.Gn = {n: Field(self.grid, self._Gn_slab[i], "Gⁿ" + n) for i, n in enumerate(names)}
        self.Gm = {n: CenterField(self.grid, "G⁻" + n) for n in names}
        self._tables = None
        self._graph = None
        self._needs_initial_update = True
        self.fused_step = bool(fused_step)
        if self.fused_step:
            u = biogeochemistry.underlying_biogeochemistry
            if not hasattr(u, "compute_tendencies_and_substep"):
                raise ValueError(f"fused_step: {type(u).__name__} has no fused tendency + substep launch")
            if getattr(biogeochemistry, "sediment", None) is not None or getattr(biogeochemistry, "particles", None) is not None:
                raise ValueError("fused_step: sediment / particles add tendencies of their own; use the three-launch path")

    # the hooks read `model.tracers` (an Oceananigans model's NamedTuple of tracer fields)
    @property
    def tracers(self):
        return self.fields

    @property
    def auxiliary_fields(self):
        return self.biogeochemistry.biogeochemical_auxiliary_fields()

    def set(self, **values):
        """`set!(model; kwargs...)` — boxmodel.jl:131-141 (scalars or one value per box)."""
        for n, v in values.items():
            if n not in self.fields:
                raise ValueError(f"name {n} not found in model.fields.")
            self.fields[n].set(v)
        self._needs_initial_update = True
        return self

    # ---- update_state!(model) — boxmodel.jl:92-110 ---------------------------------------------------------
    def _prescribed_target(self, n) -> Field:
        return self.fields[n] if n in self.fields else self.auxiliary_fields[n]

    def _apply_prescribed(self, t):
        for n, f in self.prescribed_tracers.items():
            self._prescribed_target(n).set(f(t))

    def update_state(self, compute_tendencies: bool = True):
        self._apply_prescribed(self.clock.time)
        self.biogeochemistry.update_biogeochemical_state(self)
        if compute_tendencies:
            self.compute_tendencies()

    # ---- compute_tendencies!(model) — timesteppers.jl:30-55 --------------------------------------------------
    def compute_tendencies(self):
        for n in self.Gn:  # every Gⁿ the fused launch adds into — prescribed tracers that are also biogeochemical
            f = self.forcing.get(n) if n in self.prognostic else None  # tracers included (graph mode clears the whole slab)
            if f is None:
                self.Gn[n].data.zero_()  # the per-point callable returns zero(grid), `no_func` forcing
            else:
                self.Gn[n].set(f(self.clock.time))
        self.biogeochemistry.update_tendencies(self)

    # ---- rk3_substep! + cache_previous_tendencies! in one launch — timesteppers.jl:20-28,66-93 -------------------
    def _substep(self, dt, gamma, zeta, stream=None):
        names = self.prognostic
        U = _lib.pointer_table([self.fields[n].ptr for n in names])
        Gn = _lib.pointer_table([self.Gn[n].ptr for n in names])
        Gm = _lib.pointer_table([self.Gm[n].ptr for n in names])
        cg = self.grid.c_grid()
        s = stream if stream is not None else current_stream_ptr(self.grid.device)
        rc = _lib.load().obm_rk3_substep(C.byref(cg), len(names), U, Gn, Gm, float(dt), float(gamma),
                                         0.0 if zeta is None else float(zeta), int(zeta is not None), 1, s)
        _lib.check(rc, "obm_rk3_substep")

    def _fused_stage(self, dt, gamma, zeta, stream=None):
        """One stage as one launch: Gⁿ of every tracer from the current state (+ the forcing waiting in `Gn`), the tracer
        update and the G⁻ cache."""
        forced = {n: self.Gn[n] for n in self.prognostic if self.forcing.get(n) is not None}
        self.biogeochemistry.underlying_biogeochemistry.compute_tendencies_and_substep(
            self.grid, self.fields, self.auxiliary_fields, {n: self.Gm[n] for n in self.prognostic}, dt, gamma, zeta,
            G=forced or None, accumulate=bool(forced), stream=stream)

    def time_step(self, dt: float):
        """`time_step!(model, Δt)`: Oceananigans' RK3 (or forward Euler) over the box-model methods above."""
        if self._needs_initial_update:  # what `run!` / the first `time_step!` do at iteration 0
            self.update_state(compute_tendencies=not self.fused_step)
            self._needs_initial_update = False
        stages = self.RK3 if self.timestepper == "RungeKutta3" else ((1.0, None),)
        for gamma, zeta in stages:
            self.clock.rk3_gamma, self.clock.rk3_zeta = gamma, (zeta if zeta is not None else float("nan"))
            if self.fused_step:
                for n in self.prognostic:  # the forcing the three-launch path evaluated at the end of the previous stage
                    if self.forcing.get(n) is not None:
                        self.Gn[n].set(self.forcing[n](self.clock.time))
                self._fused_stage(dt, gamma, zeta)
            else:
                self._substep(dt, gamma, zeta)
            stage_dt = dt * (gamma + (zeta or 0.0))
            self.clock.time += stage_dt
            self.clock.last_stage_dt = stage_dt
            self.update_state(compute_tendencies=not self.fused_step)
        self.clock.iteration += 1

    # ---- run!(simulation) --------------------------------------------------------------------------------------
    def run(self, dt: float, steps: int, graph: bool = False, output_every: int = 0, output_names=None,
            output: Optional[SpeedyOutput] = None, device_loop: bool = False):
        """Integrate `steps` time steps.  Returns a dict name → tensor (n_outputs, n_boxes) of snapshots taken
        every `output_every` steps (device-resident).  `output`: a `SpeedyOutput` called at iteration 0 and every
        `output_every` steps like `simulation.callbacks[:output] = Callback(SpeedyOutput(f), IterationInterval(n))`
        (eager mode), and written to disk at the end of the run.

        `device_loop=True` (with `fused_step=True`; NPZD / LOBSTER family, only PAR and T prescribed, no forcing): the
        WHOLE run is one launch — every thread integrates its box through all stages (`obm_npd_box_run`), the tabulated
        series and the snapshots never leave the device; bit-identical to `graph=True`.

        `graph=True`: prescribed tracers and forcings are tabulated for every stage of the run, uploaded once,
        and ONE captured time step is replayed `steps` times; requires a biogeochemistry whose kernel parameters do
        not depend on the clock (NPZD / LOBSTER family with prescribed or computed PAR)."""
        names = list(output_names or self.prognostic)
        nout = steps // output_every if output_every else 0
        out = {n: torch.empty((nout, self.grid.Nx), dtype=torch.float64, device=self.grid.device) for n in names}
        if output is not None and (graph or device_loop):
            raise ValueError("SpeedyOutput is an eager-mode callback; graph=True / device_loop=True return the snapshots as tensors")
        if device_loop:
            return self._run_device_loop(dt, steps, output_every, names, out)
        if not graph:
            if output is not None and output_every:
                output(self)  # IterationInterval fires at iteration 0 too
            for it in range(steps):
                self.time_step(dt)
                if output_every and (it + 1) % output_every == 0:
                    for n in names:
                        out[n][(it + 1) // output_every - 1].copy_(self.fields[n].interior.reshape(-1))
                    if output is not None:
                        output(self)
            if output is not None:
                output.flush()
            return out
        return self._run_graph(dt, steps, output_every, names, out)

    # .. graph mode ...............................................................................................
    def _tabulate(self, dt, steps):
        """Stage times of the whole run and the time series evaluated at them: row r = step·nstages + stage holds the
        values `update_state!` sees after that stage's substep."""
        stages = self.RK3 if self.timestepper == "RungeKutta3" else ((1.0, None),)
        t, times = self.clock.time, []
        for _ in range(steps):
            for gamma, zeta in stages:
                t += dt * (gamma + (zeta or 0.0))
                times.append(t)
        dev, nx = self.grid.device, self.grid.Nx

        def table(fn):
            """(rows, 1) when the series is the same for every box (one upload, broadcast on the device), else (rows, nx)."""
            vals = [fn(tt) for tt in times]
            if not any(torch.is_tensor(v) or getattr(v, "shape", ()) not in ((), (1,)) for v in vals):
                return torch.tensor([float(v) for v in vals], dtype=torch.float64).reshape(-1, 1).to(dev)
            rows = [torch.as_tensor(v, dtype=torch.float64).reshape(-1).expand(nx) for v in vals]
            return torch.stack(rows).contiguous().to(dev)

        tabs = {("prescribed", n): table(f) for n, f in self.prescribed_tracers.items()}
        tabs.update({("forcing", n): table(f) for n, f in self.forcing.items() if f is not None and n in self.prognostic})
        return times, tabs

    def _run_device_loop(self, dt, steps, output_every, names, out):
        u = self.biogeochemistry.underlying_biogeochemistry
        if not self.fused_step or not hasattr(u, "run_boxes"):
            raise ValueError("device_loop=True needs fused_step=True and a biogeochemistry with a whole-run launch (NPZD / LOBSTER family)")
        if any(f is not None for f in self.forcing.values()):
            raise ValueError("device_loop=True: forcings are not tabulated for the whole-run launch; use graph=True")
        extra = set(self.prescribed_tracers) - {"PAR", "T"}
        if extra or "PAR" not in self.prescribed_tracers:
            raise ValueError(f"device_loop=True prescribes PAR (and optionally T) from tables; got {sorted(self.prescribed_tracers)}")
        if self._needs_initial_update:
            self.update_state(compute_tendencies=False)
            self._needs_initial_update = False
        stages = self.RK3 if self.timestepper == "RungeKutta3" else ((1.0, None),)
        times, tabs = self._tabulate(dt, steps)
        if not times:
            return out
        snaps = {n: out[n] for n in names if n in self.prognostic} if output_every else None
        on_gpu = torch.device(self.grid.device).type == "cuda"
        e0, e1 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if on_gpu else (None, None)
        if on_gpu:
            e0.record()
        T_table = tabs.get(("prescribed", "T"))
        reads_T = "T" in u.required_biogeochemical_tracers()
        u.run_boxes(self.grid, self.fields, self.auxiliary_fields, {n: self.Gm[n] for n in self.prognostic}, dt, stages, steps,
                    tabs[("prescribed", "PAR")], T_table if reads_T else None, output_every, snaps)
        if T_table is not None and not reads_T:  # a prescribed series nothing reads: leave the field as the per-stage path does
            self.fields["T"].interior.reshape(-1).copy_(T_table[-1].reshape(-1).expand(self.grid.Nx))
        if on_gpu:
            e1.record()
        self.replay_events = (e0, e1)
        self.clock.time = times[-1]
        self.clock.iteration += steps
        self.clock.last_stage_dt = dt * (stages[-1][0] + (stages[-1][1] or 0.0))
        return out

    def _run_graph(self, dt, steps, output_every, names, out):
        if getattr(self.biogeochemistry.underlying_biogeochemistry, "clock_dependent_parameters", False):
            raise ValueError("graph=True needs kernel parameters that do not depend on the clock (not PISCES)")
        if callable(getattr(self.biogeochemistry.light_attenuation, "surface_PAR", None)):
            raise ValueError("graph=True: a surface PAR function of time is evaluated on the host; prescribe PAR "
                             "(prescribed_tracers={'PAR': f}) or use a constant surface PAR")
        if self._needs_initial_update:
            self.update_state(compute_tendencies=not self.fused_step)
            self._needs_initial_update = False
        stages = self.RK3 if self.timestepper == "RungeKutta3" else ((1.0, None),)
        times, tabs = self._tabulate(dt, steps)
        dev = self.grid.device
        row = torch.zeros(1, dtype=torch.long, device=dev)  # device-side cursor into the tables

        if self.fused_step:  # the first stage finds the forcing of the initial time in Gⁿ, every later one that of the row before
            for n in self.prognostic:
                if self.forcing.get(n) is not None:
                    self.Gn[n].set(self.forcing[n](self.clock.time))

        def one_step():
            for gamma, zeta in stages:
                if self.fused_step:
                    self._fused_stage(dt, gamma, zeta)
                else:
                    self._substep(dt, gamma, zeta)
                for (kind, n), tab in tabs.items():
                    if kind == "prescribed":
                        self._prescribed_target(n).interior.reshape(-1).copy_(tab.index_select(0, row).reshape(-1))
                self.biogeochemistry.update_biogeochemical_state(self)
                if not self.fused_step:
                    self._Gn_slab.zero_()
                for n in self.prognostic:
                    tab = tabs.get(("forcing", n))
                    if tab is not None:
                        self.Gn[n].interior.reshape(-1).copy_(tab.index_select(0, row).reshape(-1))
                if not self.fused_step:
                    self.biogeochemistry.update_tendencies(self)
                row.add_(1)

        # warm-up on a side stream (allocator, lazy module loads), with the state restored afterwards
        saved = [(f, f.data.clone()) for d in (self.fields, self.Gn, self.Gm, self.auxiliary_fields) for f in d.values()]
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            one_step()
        torch.cuda.current_stream(dev).wait_stream(s)
        for f, v in saved:
            f.data.copy_(v)
        row.zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            one_step()
        # the capture itself does not execute: state and cursor are still those of step 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(steps):
            g.replay()
            if output_every and (it + 1) % output_every == 0:
                for n in names:
                    out[n][(it + 1) // output_every - 1].copy_(self.fields[n].interior.reshape(-1))
        e1.record()
        self._graph, self.replay_events = g, (e0, e1)  # elapsed_time after a synchronize = the replays alone
        self.clock.time = times[-1] if times else self.clock.time
        self.clock.iteration += steps
        self.clock.last_stage_dt = dt * (stages[-1][0] + (stages[-1][1] or 0.0))
        return out

    def summary(self):
        return "Biogeochemical box model"

    def __repr__(self):
        return (f"{self.summary()}\n  Biogeochemical model: \n    └── {self.biogeochemistry.summary()}\n"
                f"  Time-stepper:\n    └── {self.timestepper}TimeStepper\n  Boxes: {self.grid.Nx}\n"
                f"  Time:\n    └── {self.clock.time} s")
