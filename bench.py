#!/usr/bin/env python
"""bench.py — BGC hot-path throughput (Gcell-updates/s) on N B200s, with per-kernel and whole-stage roofline, CPU
baselines, end-to-end (host-buffer) number, the N-GPU inventory all-reduce and a clock record.
Contract: the task statement / DESIGN.md §5.

One "step" = one Runge–Kutta stage of the biogeochemistry for the whole grid, i.e. exactly what Oceananigans triggers
per stage through the plugin hooks:
    update_biogeochemical_state!(bgc, model)   → negative scaling (+ Ω for PISCES), PAR scan (+ zₑᵤ, PAR̄), sediment
    update_tendencies!(bgc, model)             → fused tendencies of every tracer, Gⁿ += …, sediment ↔ tracer fluxes
A "cell-update" = all of that for one grid cell.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference] [--scaling strong|weak]

Under torchrun (N > 1) the ONE grid of the workload is split into N x–y slabs along y (BASELINE configs[3]: "1024×1024×128
x-y slab-sharded across 1/2/4/8 B200") — `"scaling": "strong"`; there is no data-path collective (every kernel is
pointwise or column-local), time = max over ranks.  The same run also reports, as sub-records: the weak-scaling number
(one full-size grid per GPU), the tracer-inventory diagnostic executed every stage (obm_inventory + the path's only
collective, an all-reduce of ≤ 8 doubles over NCCL) with its own time and a check against the oracle's serial sum, and
the end-to-end leg with the PCIe peak measured while all N ranks copy concurrently.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


# --------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# --------------------------------------------------------------------------------------------------
def eady_z_faces(Nz=64, Lz=140.0, refinement=1.8, stretching=3.0):
    """The stretched vertical grid of the reference's Eady case, paper/figures/eady.jl:11-27 (k = 1 … Nz + 1)."""
    def z(k):
        h = (k - 1) / Nz
        zeta0 = 1 + (h - 1) / refinement
        sigma = (1 - np.exp(-stretching * h)) / (1 - np.exp(-stretching))
        return Lz * (zeta0 * sigma - 1)
    return z


def workload_table():
    return {
        # name: (description, builder)
        "lobster_c3": ("LOBSTER + carbonates + O2 with SimpleMultiG sediment bottom boundary and sinking sPOM / bPOM, 3-D Eady grid "
                       "512x512x64 on the stretched z of paper/figures/eady.jl (BASELINE configs[2])",
                       dict(model="lobster", size=(512, 512, 64), extent=(1000.0, 1000.0, 140.0), sediment=True, z="eady")),
        "lobster_c2": ("LOBSTER + carbonates + O2, column ensemble 4096 columns x 64 levels (BASELINE configs[1])",
                       dict(model="lobster", size=(4096, 1, 64), extent=(4096.0, 1.0, 200.0))),
        "npzd_c1": ("NPZD + TwoBandPAR on the README grid 160x1x32 (BASELINE configs[0])",
                    dict(model="npzd", size=(160, 1, 32), extent=(10e3, 1.0, 500.0))),
        "pisces_c4": ("PISCES + 3-band PAR + calcite saturation, 1024x1024x128 (BASELINE configs[3], the headline)",
                      dict(model="pisces", size=(1024, 1024, 128), extent=(1024e3, 1024e3, 400.0))),
        "carbon_c5": ("CarbonChemistry pH solve sweep over 1e8 synthetic (T, S, DIC, Alk) cells (BASELINE configs[4])",
                      dict(model="carbon", n=100_000_000)),
    }


def default_workload():
    return "pisces_c4"


def shards(name, world, scaling):
    """How the workload is split over `world` ranks: ("strong", rows per rank) when its grid divides along y, else
    ("weak", None) — independent replicas (column ensembles and the flat sweep have no y to split)."""
    cfg = workload_table()[name][1]
    if scaling == "weak" or world == 1:
        return ("weak" if world > 1 else scaling), None
    if cfg["model"] == "carbon" or cfg["size"][1] % world or cfg["size"][1] // world < 1 or cfg["size"][1] == 1:
        return "weak", None
    return "strong", cfg["size"][1] // world


def config_dict(name, world, scaling):
    """The `config` of the JSON line — built by this one function for BOTH arms (`--impl b200` and `--impl reference`),
    so the driver's same-config check compares like with like."""
    desc, cfg = workload_table()[name]
    mode, rows = shards(name, world, scaling)
    if cfg["model"] == "carbon":
        grid = f"flat n = {cfg['n']}"
        small = False
    else:
        Nx, Ny, Nz = cfg["size"]
        grid = f"{Nx}x{Ny}x{Nz}"
        small = Nx * Ny * Nz * 640 <= 4e8
    return {"workload": name, "description": desc, "grid": grid, "scaling": mode,
            "parallelism": (f"one grid, {world} x-y slabs of {rows} rows along y" if mode == "strong" and world > 1
                            else (f"{world} independent replicas of the grid" if world > 1 else "1 GPU")) + "; no data-path collective",
            "l2": ("working set fits L2: evicted by a 256 MB write before every step (not timed); time = sum of per-step event pairs"
                   if small else "inputs larger than L2 (no flush needed)"),
            "state": "static synthetic fields (splitmix64, seed 20260117), all positive: the negative-scaling pass reads every "
                     "scaled tracer and rewrites none",
            "carbonate_solve": ("cold every step (no stored [H+]: warm start disabled for the static state): carbonate-alkalinity "
                                "quadratic + 3 Newton steps in ln[H+] in FP32, then FP64 Newton steps (one for sea water) with the "
                                "second-order error term removed" if cfg["model"] in ("pisces", "carbon") else None)}


class Workload:
    """Builds the model state on `device` (synthetic fields) and exposes step().
    `rows = (j0, ny, Ny_global)`: this rank's y-slab of the global grid (strong scaling)."""

    def __init__(self, name, device, scale=1.0, rows=None):
        import oceanbiome_b200 as ob
        from oceanbiome_b200 import synthetic
        self.ob, self.name, self.device = ob, name, device
        desc, cfg = workload_table()[name]
        self.description, self.cfg = desc, cfg
        self.kind = cfg["model"]
        self.marks = None
        if self.kind == "carbon":
            self._build_carbon(int(cfg["n"] * scale))
            return
        Nx, Ny, Nz = cfg["size"]
        if scale != 1.0:
            Ny = max(1, int(Ny * scale))
        Lx, Ly, Lz = cfg["extent"]
        fill_rows = None
        if rows is not None:
            j0, ny, nyg = rows
            Ly, Ny = Ly * ny / Ny, ny
            fill_rows = (j0, nyg)
        topo = ("Periodic", "Periodic" if Ny > 1 else "Flat", "Bounded")
        size = (Nx, Ny, Nz) if Ny > 1 else (Nx, Nz)
        z = eady_z_faces(Nz, Lz) if cfg.get("z") == "eady" else (-Lz, 0.0)
        kw = dict(x=(0.0, Lx), z=z) if Ny == 1 else dict(x=(0.0, Lx), y=(0.0, Ly), z=z)
        self.grid = ob.RectilinearGrid(size=size, topology=topo, device=device, **kw)
        sediment = None
        if cfg.get("sediment"):
            # configs[2] as named: SimpleMultiG sediment under sinking sPOM / bPOM (detritus.jl:35-36 speeds), the
            # sediment's own AB2 stepper and its first-order upwind bottom flux (test_sediments.jl:37-80 setup)
            sediment = ob.SimpleMultiGSediment(self.grid)
        if self.kind == "lobster":
            self.bgc = ob.LOBSTER(self.grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen(), sediment=sediment,
                                  scale_negatives=True, surface_photosynthetically_active_radiation=100.0)
            ranges = synthetic.lobster_range
        elif self.kind == "npzd":
            self.bgc = ob.NPZD(self.grid, scale_negatives=True, surface_photosynthetically_active_radiation=100.0)
            ranges = lambda n: synthetic.RANGES_NPZD[n]  # noqa: E731
        elif self.kind == "pisces":
            self.bgc = ob.PISCES(self.grid, scale_negatives=True, surface_photosynthetically_active_radiation=100.0)
            # the synthetic state does not change between steps: with the Newton warm start on, every Ω solve after the
            # first would converge in one iteration, which no simulation sees ⇒ the bench always solves cold
            self.bgc.underlying_biogeochemistry.warm_start_carbonate_solve = False
            ranges = ob.pisces.synthetic_range
        self.model = ob.BiogeochemicalModel(self.grid, self.bgc)
        for n, f in self.model.tracers.items():
            lo, hi, log = ranges(n)
            synthetic.fill_torch(f, n, lo, hi, log, rows=fill_rows)
        if self.kind == "pisces":
            ob.pisces.fill_synthetic_auxiliary(self.bgc, self.model)
        if sediment is not None:
            for n, f in sediment.fields.items():  # pools Ns, Nf, Nr ∈ [1e-2, 10] (log), SURVEY §8d C3
                synthetic.fill_torch(f, "sed" + n, 1e-2, 10.0, True)
            self.model.clock.last_stage_dt = 60.0  # the sediment hook steps its pools with Δt of the last stage
        self.ranges = ranges
        self.cells = self.grid.ncells
        self.nG = sum(1 for n in self.model.tracers if n not in ("T", "S"))
        mods = self.bgc.modifiers if isinstance(self.bgc.modifiers, tuple) else (self.bgc.modifiers,)
        self.groups = [(m.tracers, m.scalefactors) for m in mods]
        scaled = []
        for tn, _ in self.groups:
            scaled += [t for t in tn if t not in scaled]
        self.stage_kernels = self._stage_kernels(len(scaled), sediment is not None)
        self.tendency_bytes_per_cell = self.stage_kernels["tendencies"][1]
        self.stage_bytes_per_cell = sum(b for _, b, _ in self.stage_kernels.values())
        self.step_kernels = [k for k, _, _ in self.stage_kernels.values()]

    def _stage_kernels(self, nscaled, sediment):
        """marker label → (kernel, ALGORITHMIC bytes per cell — SURVEY §8d's per-unit figures —, what bounds it)."""
        if self.kind == "pisces":
            # 20 distinct scaled tracers (carbon 9, iron + 5, phosphate + 1, silicon + 3, nitrogen + 2) + Alk + T + S read, Ω written
            # = 192 B (r01 / r02 documents said 200: one field too many)
            return {"modifiers": ("scale_negative_calcite_kernel", 8 * nscaled + 8 + 16 + 8, "issue / FP64 (Ω solve); HBM roof shown"),
                    "light": ("par_multiband_kernel<3,DIAG>", 16 + 32, "issue / FP64 (1 log + 6 exp per cell); HBM roof shown"),
                    "tendencies": ("pisces_tendency_kernel", 256 + 16 * 24, "latency of dependent FP64 chains; HBM roof shown")}
        npd = {"lobster": 8 * (7 + 1) + 16 * 10, "npzd": 8 * (5 + 1) + 16 * 4}[self.kind]
        k = {"modifiers": ("scale_negative_kernel", 8 * nscaled, "hbm"),
             "light": ("par_twoband_kernel", 16, "issue / FP64 (2 pow + 2 exp per cell); HBM roof shown"),
             "tendencies": ("npd_tendency_kernel", npd, "hbm")}
        if sediment:  # per COLUMN ≈ 200 B = 200 / Nz per cell
            k["sediment"] = ("sediment_state_kernel", 200.0 / self.grid.Nz, "latency (Nx·Ny threads)")
            k["sediment_tendencies"] = ("sediment_tendency_kernel", 100.0 / self.grid.Nz, "latency (Nx·Ny threads)")
        return k

    def _build_carbon(self, n):
        from oceanbiome_b200 import synthetic
        dev = self.device
        self.n = n
        self.cells = n
        u = lambda name: synthetic.uniform_torch(synthetic.field_id(name), 0, n, dev)  # noqa: E731
        self.T = -2.0 + 37.0 * u("T")
        self.S = 20.0 + 20.0 * u("S")
        self.DIC = 1800.0 + 600.0 * u("DIC")
        lo = torch.maximum(self.DIC * 1.02, torch.full_like(self.DIC, 2000.0))
        self.Alk = lo + (2600.0 - lo) * u("Alk")
        self.out = torch.empty_like(self.DIC)
        self.cc = self.ob.CarbonChemistry(newton_iterations=12)
        self.tendency_bytes_per_cell = self.stage_bytes_per_cell = 40
        self.stage_kernels = {"tendencies": ("carbon_sweep_kernel", 40, "FP64 pipe; HBM roof shown")}
        self.step_kernels = ["carbon_sweep_kernel"]
        self.groups = []

    def capture(self):
        """One stage (all hooks) as a CUDA graph on the current stream; parameters evaluated on the host (day length,
        surface PAR) are frozen at their capture-time values, which is what a static synthetic bench state has anyway."""
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.step()
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.step()
        return self.graph

    # ---- one step of the hot path, inputs resident in HBM -------------------------------------------
    def step(self, marks=None):
        """`marks`: list that receives (label, CUDA event) after every launch group of the stage (first entry: start)."""
        def mark(label):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((label, e))
        if self.kind == "carbon":
            if marks is not None:
                mark("start")
            self.cc(DIC=self.DIC, T=self.T, S=self.S, Alk=self.Alk, output="pHᶠ", out=self.out)
            if marks is not None:
                mark("tendencies")
            return 1
        m = self.model
        bgc = m.biogeochemistry
        if marks is not None:
            mark("start")
            bgc.stage_marker = mark
        bgc.update_biogeochemical_state(m)
        bgc.update_tendencies(m)
        bgc.stage_marker = None
        return len(self.step_kernels)


def kernel_times(step_marks, table):
    """Mean duration (ms) of every launch group over the steps, from the marker events recorded inside the timed region."""
    acc = {}
    for marks in step_marks:
        for (_, e0), (label, e1) in zip(marks[:-1], marks[1:]):
            if label in table:
                acc.setdefault(label, []).append(e0.elapsed_time(e1))
    return {k: float(np.mean(v)) for k, v in acc.items()}


# --------------------------------------------------------------------------------------------------
# end-to-end leg: host (pinned) buffers in, host buffers out, through the public API
# --------------------------------------------------------------------------------------------------
class HostStage:
    """The call a user with HOST arrays makes: tracers live in pinned host memory; every step copies them to
    the device, runs the stage through the plugin hooks and reads every tendency back — as a 3-stream slab
    pipeline (oceanbiome_b200.host_stage.HostStagedStage).  The flat carbonate sweep copies its 4 inputs in and
    its output back."""

    def __init__(self, w: Workload, copy_engine: str = "dma", nslabs: int = 0):
        from oceanbiome_b200.host_stage import HostStagedStage, bind_to_gpu_numa_node
        self.w = w
        # first-touch the pinned buffers on the NUMA node the GPU hangs off (no-op on single-node hosts)
        self.numa_bound = bind_to_gpu_numa_node(w.device)
        if w.kind == "carbon":
            self.h_in = [torch.empty(w.n, dtype=torch.float64).pin_memory() for _ in range(4)]
            for h, d in zip(self.h_in, (w.T, w.S, w.DIC, w.Alk)):
                h.copy_(d)
            self.h_out = torch.empty(w.n, dtype=torch.float64).pin_memory()
            self.h2d_bytes = 4 * 8 * w.n
            self.d2h_bytes = 8 * w.n
            return
        self.stage = HostStagedStage(w.model, nslabs=nslabs or ((16 if w.grid.Ny >= 256 else 8) if w.grid.Ny >= 64 else 1),
                                     copy_engine=copy_engine)
        self.stage.upload_from_device()
        self.h2d_bytes, self.d2h_bytes = self.stage.h2d_bytes, self.stage.d2h_bytes

    def step(self):
        w = self.w
        if w.kind == "carbon":
            for h, d in zip(self.h_in, (w.T, w.S, w.DIC, w.Alk)):
                d.copy_(h, non_blocking=True)
            w.step()
            self.h_out.copy_(w.out, non_blocking=True)
            return
        self.stage.step()


def pcie_peak_gbs(device, nbytes=1 << 30, reps=3, barrier=None):
    """Measured host<->device rate of this box with both directions busy (pinned memory, one large contiguous copy per
    direction on its own stream): the roofline of the end-to-end leg, which moves every tracer in and every tendency
    out.  GB/s per direction.  `barrier`: called after the buffers are pinned and warmed (pinning 2 GiB takes a rank-
    dependent fraction of a second — without it the ranks' timed copies of a multi-GPU run would not overlap and the
    "concurrent" peak would be each rank's solo rate)."""
    n = nbytes // 8
    h_in, h_out = (torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2))
    d_in = torch.empty(n, dtype=torch.float64, device=device)
    d_out = torch.zeros(n, dtype=torch.float64, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)
    cur = torch.cuda.current_stream(device)

    def run(r):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_stream(cur); s2.wait_stream(cur)
        for _ in range(r):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
        b.record()
        torch.cuda.synchronize(device)
        return r * n * 8 / (a.elapsed_time(b) * 1e-3) / 1e9

    run(1)
    if barrier is not None:
        barrier()
    return run(reps)


def host_memcpy_gbs(nbytes=1 << 29, reps=3):
    """Host DRAM copy rate of one core (numpy memcpy, read + write bytes): with all ranks running it at once it shows
    what the shared host memory system leaves each rank."""
    a = np.ones(nbytes // 8)
    b = np.empty_like(a)
    np.copyto(b, a)
    t0 = time.perf_counter()
    for _ in range(reps):
        np.copyto(b, a)
    return 2 * reps * nbytes / (time.perf_counter() - t0) / 1e9


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        if self._run_nvml():
            return
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def _run_nvml(self):
        """Same quantities through NVML (≈ 5 ms per sample instead of ≈ 100 ms per nvidia-smi process)."""
        try:
            import pynvml as N
            import torch
            N.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
        except Exception:
            return False
        act = lambda bit: "Active" if bit else "Not Active"  # noqa: E731
        first = True
        while first or not self.stop_flag:  # at least one sample, however short the timed region
            first = False
            try:
                r = reasons(h)
                self.samples.append([str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx),
                                     str(N.nvmlDeviceGetPowerUsage(h) / 1000.0), act(r & 0x8), act(r & 0x40), act(r & 0x20),
                                     act(r & 0x4)])
            except Exception:
                pass
            time.sleep(0.01)
        return True

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples),
                "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, cells):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch from the committed
    `ncu --set full` capture (profiles/traffic.json), scaled to this run's cell count → (bytes | None, note)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        d = json.load(open(p)).get(workload)
        if d:
            return d["bytes_per_launch"] * cells / d["cells"], f'{d["source"]}: {d["note"]}'
    return None, None


def fp64_roofline(workload, cells, kernel_ms):
    """The dominant kernel against the OTHER roof BASELINE.json's metric names: FP64-pipe instructions per cell counted by
    ncu in the committed capture (profiles/traffic.json) × this run's cells ÷ this run's kernel time, against the DFMA
    rate measured on this pool's B200s.  None when the capture holds no count for this workload."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    d = t.get(workload) or {}
    if "fp64_instr_per_cell" not in d or "fp64_peak_instr_per_s" not in t:
        return None
    achieved = d["fp64_instr_per_cell"] * cells / (kernel_ms * 1e-3) / 1e12
    peak = t["fp64_peak_instr_per_s"] / 1e12
    return {"instr_per_cell": d["fp64_instr_per_cell"], "achieved": achieved, "peak": peak, "unit": "T FP64 instr/s",
            "frac": achieved / peak, "source": d["fp64_source"], "peak_source": t["fp64_peak_source"]}


# --------------------------------------------------------------------------------------------------
# CPU legs (the oracle = port of the reference algorithm and launch structure)
# --------------------------------------------------------------------------------------------------
def cpu_sample(name, threads, budget_cells):
    """Time the oracle (one pass per tracer, reference solver, serial-in-z PAR, one pass per group) on a
    bounded sub-volume of the same workload.  Returns (cells/s, sample description, seconds)."""
    import pyoracle
    from oceanbiome_b200 import synthetic
    import oceanbiome_b200 as ob
    pyoracle.build()
    pyoracle.set_threads(threads)
    desc, cfg = workload_table()[name]
    if cfg["model"] == "carbon":
        n = int(budget_cells)
        u = lambda nm: synthetic.uniform_numpy(synthetic.field_id(nm), 0, n)  # noqa: E731
        T, S, DIC = -2.0 + 37.0 * u("T"), 20.0 + 20.0 * u("S"), 1800.0 + 600.0 * u("DIC")
        lo = np.maximum(DIC * 1.02, 2000.0)
        Alk = lo + (2600.0 - lo) * u("Alk")
        t0 = time.perf_counter()
        pyoracle.carbon_chemistry_sweep(T, S, DIC, Alk, output=2)
        dt = time.perf_counter() - t0
        return n / dt, f"first {n} cells of the sweep ({n / cfg['n']:.4f} of it), reference damped Newton (atol 1e-20, max 100 iterations)", dt
    Nx, Ny, Nz = cfg["size"]
    ny = max(1, min(Ny, int(budget_cells // (Nx * Nz))))
    nx = Nx if ny >= 1 and Nx * Nz <= budget_cells else max(1, int(budget_cells // Nz))
    topo = ("Periodic", "Periodic" if ny > 1 else "Flat", "Bounded")
    size = (nx, ny, Nz) if ny > 1 else (nx, Nz)
    Lx, Ly, Lz = cfg["extent"]
    z = eady_z_faces(Nz, Lz) if cfg.get("z") == "eady" else (-Lz, 0.0)
    kw = dict(x=(0.0, Lx * nx / Nx), z=z) if ny == 1 else dict(x=(0.0, Lx * nx / Nx), y=(0.0, Ly * ny / Ny), z=z)
    grid = ob.RectilinearGrid(size=size, topology=topo, device="cpu", **kw)
    og = pyoracle.Grid.like(grid)
    if cfg["model"] in ("lobster", "npzd"):
        dt = npd_oracle_stage_seconds(pyoracle, grid, og, cfg)
    else:
        dt, grid = pisces_oracle_stage_seconds(pyoracle, grid, og)
    frac = grid.ncells / (Nx * Ny * Nz)
    return grid.ncells / dt, (f"{grid.Nx}x{grid.Ny}x{grid.Nz} sub-volume of the workload grid (1/{1 / frac:.0f} of it), reference launch "
                             "structure (one pass per tracer / band / group, reference solver)"), dt


def npd_oracle_stage_seconds(pyoracle, grid, og, cfg):
    """CPU-baseline leg for the LOBSTER / NPZD workloads: one stage in the reference's launch structure (one scaling pass
    per group, serial-in-z two-band PAR, the sediment's column passes, one tendency pass per tracer)."""
    import oceanbiome_b200 as ob
    from oceanbiome_b200 import synthetic
    if cfg["model"] == "lobster":
        bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen(), scale_negatives=True)
        rng = synthetic.lobster_range
    else:
        bgc = ob.NPZD(grid, scale_negatives=True)
        rng = lambda n: synthetic.RANGES_NPZD[n]  # noqa: E731
    u = bgc.underlying_biogeochemistry
    names = u.required_biogeochemical_tracers()
    host = {n: synthetic.fill_numpy(np.zeros(og.parent_shape), og, n, *rng(n)) for n in names}
    mods = bgc.modifiers if isinstance(bgc.modifiers, tuple) else (bgc.modifiers,)
    groups = [(m.tracers, m.scalefactors) for m in mods]
    snames = []
    for tn, _ in groups:
        snames += [t for t in tn if t not in snames]
    cgroups = pyoracle.make_groups(snames, groups)
    G = [np.zeros(og.parent_shape) for _ in names]
    PAR = np.zeros(og.parent_shape)
    sed, keep = None, None
    if cfg.get("sediment"):
        sed = ob.SimpleMultiGSediment(grid)
        pools = {n: synthetic.fill_numpy(np.zeros(og.plane_shape), og, "sed" + n, 1e-2, 10.0, True) for n in sed.fields}
        plane = lambda: np.zeros(og.plane_shape)  # noqa: E731
        wf = {}
        for n in sed.biogeochemistry.sinking_fluxes():
            w = np.zeros((og.parent_shape[0] + 1,) + og.parent_shape[1:])
            w[og.Hz:og.Hz + og.Nz] = float(bgc.biogeochemical_drift_velocity(n))
            wf[n] = w
        sink = list(sed.biogeochemistry.sinking_fluxes())
        keep = dict(pools=[pools[n] for n in sed.fields], Gn=[plane() for _ in sed.fields], Gm=[plane() for _ in sed.fields],
                    tracked=[plane() for _ in sed.tracked_fields], sinking_w=[wf[n] for n in sink])  # referenced until the end
        sf = pyoracle.sediment_fields(
            NO3=host["NO₃"], NH4=host["NH₄"], O2=host["O₂"], sinking=[host[n] for n in sink], sinking_w=keep["sinking_w"],
            pools=keep["pools"], Gn=keep["Gn"], Gm=keep["Gm"], tracked=keep["tracked"],
            G_coupled=[G[names.index(n)] if n in names else None for n in sed.biogeochemistry.coupled_tracers()])
        sp = sed.c_params()
    t0 = time.perf_counter()
    pyoracle.scale_negative_tracers(og, [host[n] for n in snames], cgroups)
    pyoracle.par_twoband(og, bgc.light_attenuation.c_params(), host["P"], 100.0, PAR)
    if sed is not None:
        pyoracle.sediment_update_state(og, sp, sf, 60.0)
    pyoracle.npd_tendencies(og, u.c_params(), [host[n] for n in names], PAR, G=G, accumulate=True)
    if sed is not None:
        pyoracle.sediment_update_tendencies(og, sp, sf)
    dt = time.perf_counter() - t0
    del keep
    return dt


def pisces_oracle_stage_seconds(pyoracle, grid, og):
    """CPU-baseline leg of bench.py for the PISCES workload: one stage in the reference's launch structure
    (5 scaling passes, 3 PAR passes, zₑᵤ, PAR̄, Ω with the reference's damped Newton, 24 tendency passes) on the
    host twin `og` of `grid`.  TEST/BENCH INFRASTRUCTURE: the oracle is passed in by the caller."""
    import time as _time

    from oceanbiome_b200 import synthetic
    from oceanbiome_b200.pisces import PISCES, TRACERS, DepthDependantSinkingSpeed, synthetic_range
    bgc = PISCES(grid, scale_negatives=True)
    u = bgc.underlying_biogeochemistry
    host = {n: synthetic.fill_numpy(np.zeros(og.parent_shape), og, n, *synthetic_range(n)) for n in TRACERS}
    zmxl = synthetic.fill_numpy(np.zeros(og.plane_shape), og, "zₘₓₗ", -150.0, -10.0)
    kappa = synthetic.fill_numpy(np.zeros(og.plane_shape), og, "κ̄", 1e-4, 1e-2, True)
    wPOC = np.ascontiguousarray(u.sinking_velocities["POC"].data.numpy())
    u.mixed_layer_depth.data.copy_(torch.from_numpy(zmxl))
    u.euphotic_depth.data.fill_(-60.0)
    wGOC = np.ascontiguousarray(DepthDependantSinkingSpeed().face_field(grid, u.mixed_layer_depth, u.euphotic_depth).data.numpy())
    groups = [(m.tracers, m.scalefactors) for m in bgc.modifiers]
    snames = []
    for tn, _ in groups:
        snames += [t for t in tn if t not in snames]
    cgroups = pyoracle.make_groups(snames, groups)
    la = bgc.light_attenuation
    G = [np.zeros(og.parent_shape) if n < 24 else None for n in range(26)]
    t0 = _time.perf_counter()
    pyoracle.scale_negative_tracers(og, [host[n] for n in snames], cgroups)
    bands, total = pyoracle.par_multiband(og, la.c_params(), host["PChl"], host["DChl"], 1.0, 100.0)
    zeu = pyoracle.euphotic_depth(og, total)
    mean = pyoracle.mixed_layer_mean(og, zmxl, total)
    Om = pyoracle.calcite_saturation(og, host["T"], host["S"], host["DIC"], host["Alk"], host["Si"])
    aux = {"PAR1": bands[0], "PAR2": bands[1], "PAR3": bands[2], "PAR": total, "Omega": Om, "wPOC": wPOC, "wGOC": wGOC,
           "mixed_layer_depth_xy": zmxl, "euphotic_depth_xy": zeu, "mean_mixed_layer_vertical_diffusivity_xy": kappa,
           "mean_mixed_layer_light_xy": mean}
    pyoracle.pisces_tendencies(og, u.c_params(0.0), [host[n] for n in TRACERS], aux, G=G, accumulate=True)
    return _time.perf_counter() - t0, grid


def reference_budget(name):
    """Cells of the bounded sample the reference arm times per step: ≥ 1/64 of the workload."""
    cfg = workload_table()[name][1]
    if cfg["model"] == "carbon":
        return cfg["n"] // 32
    Nx, Ny, Nz = cfg["size"]
    total = Nx * Ny * Nz
    if cfg["model"] == "pisces":
        return max(total // 64, Nx * Nz)
    return max(total // 8, Nx * Nz) if total > 4_000_000 else total


def workload_cells(name):
    cfg = workload_table()[name][1]
    return cfg["n"] if cfg["model"] == "carbon" else int(np.prod(cfg["size"]))


def fused_cpu_sample(name, threads, budget_cells):
    """The second CPU column of BASELINE.md §3: the GPU kernels' FUSED algorithm (one pass over the cells, every shared
    sub-model evaluated once, fixed-iteration Newton) compiled for the host — separates the algorithmic gain from the
    hardware gain.  Needs bench_ref/libobm_fused_host.so (bench_ref/Makefile); None when it is not built."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "bench_ref"))
        import fused_host
    except Exception:
        return None
    return fused_host.sample(name, threads, budget_cells, workload_table()[name][1])


def run_reference_arm(args, name):
    """--impl reference: the reference's CPU implementation of the path (here: its C restatement — Julia
    is not in the image) with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    budget = reference_budget(name)
    vals, secs = [], []
    for s in range(args.warmup + args.steps):
        v, sample, dt = cpu_sample(name, threads, budget)
        if s >= args.warmup:
            vals.append(v)
            secs.append(dt)
    value = float(np.mean(vals)) / 1e9
    world = max(1, args.gpus)
    line = {
        "impl": "reference", "metric": "BGC tendency Gcell-updates/s", "value": value, "unit": "Gcell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # the time ONE stage of the whole workload grid would take at the sampled rate (the sample's own time is below)
        "ms_per_step": 1e3 * workload_cells(name) / (value * 1e9), "higher_is_better": True,
        "scaling": config_dict(name, world, args.scaling)["scaling"], "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(name, world, args.scaling),
        "cpu_baseline": {"value": value, "unit": "Gcell-updates/s", "cores": threads, "kind": "port", "sample": sample,
                         "sample_ms": 1e3 * float(np.mean(secs)),
                         "note": "C restatement of the reference algorithm and launch structure (oracle/): Julia is not in the image"},
        "e2e": {"value": value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def inventory_check(w, diag, rank, world, rows_info):
    """The inventory path (obm_inventory → all-reduce) against the oracle's serial sum, on a bounded set of cells: the
    first R rows of every rank's slab.  The device side is the real thing — kernel on the sub-range, NCCL all-reduce —
    and rank 0 regenerates exactly those rows of the global synthetic field on the host for the oracle."""
    import torch.distributed as dist
    g = w.grid
    R = min(2, g.Ny)
    with g.restrict(0, R):
        got = diag.local().clone()
    if world > 1:
        dist.all_reduce(got, op=dist.ReduceOp.SUM)
    got = got.cpu().numpy()
    if rank != 0:
        return None
    import pyoracle
    from oceanbiome_b200 import synthetic
    import oceanbiome_b200 as ob
    names = diag.names
    cg = pyoracle.make_groups(names, diag.groups)
    want = np.zeros(len(diag.groups))
    mag = np.zeros(len(diag.groups))
    ny_local = g.Ny
    nyg = rows_info[2] if rows_info else g.Ny
    for r in range(world if rows_info else 1):
        sub = ob.RectilinearGrid(size=(g.Nx, R, g.Nz), x=(0.0, g.Lx), y=(0.0, g.dy * R), z=g.zf, device="cpu") if g.Ny > 1 else None
        if sub is None:
            return {"skipped": "flat y"}
        og = pyoracle.Grid.like(sub)
        fields = [synthetic.fill_numpy(np.zeros(og.parent_shape), og, n, *w.ranges(n), rows=(r * ny_local, nyg)) for n in names]
        vol = np.zeros(og.parent_shape)
        og.interior(vol)[...] = (g.dz * g.dx * g.dy).reshape(-1, 1, 1)
        want += pyoracle.inventory(og, fields, cg, cell_volume=vol)
        mag += pyoracle.inventory(og, [np.abs(f) for f in fields],
                                  pyoracle.make_groups(names, [(tn, tuple(abs(x) for x in sf)) for tn, sf in diag.groups]), cell_volume=vol)
    if world > 1 and not rows_info:  # independent replicas of the same grid: the all-reduce adds `world` equal inventories
        want, mag = want * world, mag * world
    err = float(np.max(np.abs(got - want) / mag))
    return {"cells": int(g.Nx * R * g.Nz * (world if rows_info else 1)), "rows_per_rank": R, "max_err_over_sum_abs_terms": err,
            "tolerance": 1e-12, "ok": bool(err <= 1e-12),
            "how": "obm_inventory on the first rows of every slab + all-reduce vs the oracle's serial (long double) sum of the same cells"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=list(workload_table()))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink Ny (or n) for quick checks; not a bench number")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): the workload's ONE grid split into N y-slabs (BASELINE configs[3]); weak: one full grid per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling sub-record of a multi-GPU run")
    ap.add_argument("--e2e-balance-threshold", type=float, default=1.05,
                    help="re-share the rows when the slowest rank's equal-slab stage takes more than this × the fastest's")
    ap.add_argument("--no-e2e-balance", action="store_true",
                    help="multi-GPU e2e leg: keep equal slabs (default: re-share the rows by each rank's measured host-link rate)")
    ap.add_argument("--no-inventory", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the stage as one CUDA graph (auto: when the working set fits L2, i.e. the stage is launch-bound)")
    ap.add_argument("--e2e-slabs", type=int, default=0, help="x-y slabs of the e2e pipeline (0: default)")
    ap.add_argument("--copy-engine", default="dma", choices=["sm", "dma", "sm_h2d", "sm_d2h"],
                    help="host<->device slab copies of the e2e leg: persistent copy kernel (sm) or cudaMemcpy2DAsync (dma)")
    args = ap.parse_args()
    name = args.workload or default_workload()
    if args.impl == "reference":
        run_reference_arm(args, name)
        return

    # stdout carries exactly ONE line — the JSON — so everything libraries print there while the job runs (NCCL's
    # "NCCL version …" banner at communicator creation, for one) is sent to stderr until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import oceanbiome_b200 as ob
    from oceanbiome_b200.distributed import InventoryDiagnostic, init_distributed
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: oceanbiome.jl_b200 has no CPU fallback")
    ob.load_library()
    rank, world, device = init_distributed()
    if args.warmup < 3:
        args.warmup = 3

    mode, nrows = shards(name, world, args.scaling)
    cfg = workload_table()[name][1]
    rows_info = None
    if mode == "strong" and world > 1:
        nyg = max(world, int(cfg["size"][1] * args.scale) // world * world)
        rows_info = (rank * (nyg // world), nyg // world, nyg)
    w = Workload(name, device, args.scale, rows=rows_info)
    sampler = ClockSampler(device.index or 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    def timed_steps(wl, steps, use_graph, small, flush, marks_out=None, extra=None):
        """K steps between two events on the launching stream (per-step pairs around the L2 flush for small working
        sets); `extra()` runs after every step inside the region.  → (ms on this rank, launches)."""
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        t0.record()
        for s in range(steps):
            if small:
                flush.fill_(float(s))
                sev[s][0].record()
            if use_graph:
                wl.graph.replay()
                launches += len(wl.step_kernels)
            else:
                marks = [] if marks_out is not None else None
                launches += wl.step(marks)
                if marks_out is not None:
                    marks_out.append(marks)
            if extra is not None:
                launches += extra()
            if small:
                sev[s][1].record()
        t1.record()
        barrier()
        return (sum(a.elapsed_time(b) for a, b in sev) if small else t0.elapsed_time(t1)), launches

    for _ in range(args.warmup):
        w.step()
    barrier()
    # A working set that fits the 126 MB L2 (configs C1, C2) is (a) evicted before every step by a 256 MB write that
    # is left out of the timing, and (b) launch-bound: its stage — three tiny launches — is replayed as ONE CUDA
    # graph, which removes the host's per-launch cost (≈ 35 µs per hook call from Python) the way the box-model driver
    # does.  Large workloads keep one event pair around all K steps.
    small = w.cells * w.tendency_bytes_per_cell <= 4e8
    use_graph = args.graph == "on" or (args.graph == "auto" and small and w.kind != "carbon")
    if use_graph:
        w.capture()
    flush = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device=device) if small else None

    # ---- timed region A: K stages, CUDA events on the launching stream, every launch group marked -----------------
    step_marks = []
    sampler.start()
    ms, launches = timed_steps(w, args.steps, use_graph, small, flush, marks_out=None if use_graph else step_marks)
    sampler.stop_flag = True
    if use_graph:  # per-kernel times of a graphed stage: a few eager steps with the markers (they include launch latency)
        for s in range(args.steps):
            flush.fill_(float(s)) if small else None
            marks = []
            w.step(marks)
            step_marks.append(marks)
        torch.cuda.synchronize(device)
    ktimes = kernel_times(step_marks, w.stage_kernels)
    kernel_ms = ktimes["tendencies"]
    ms = max_over_ranks(ms)
    global_cells = w.cells * world
    value = global_cells * args.steps / (ms * 1e-3) / 1e9

    # ---- timed region B: the same stages with the conservation diagnostic every stage ------------------------------
    inventory = None
    if not args.no_inventory and w.groups and w.kind != "carbon":
        diag = InventoryDiagnostic(w.grid, w.model.tracers, w.groups)
        inv_ev = []

        def run_inventory():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            diag()  # obm_inventory (2 launches) + all-reduce of len(groups) doubles
            b.record()
            inv_ev.append((a, b))
            return 2
        for _ in range(2):
            w.step(); run_inventory()
        barrier()
        inv_ev.clear()
        ms_b, _ = timed_steps(w, args.steps, False, small, flush, extra=run_inventory)
        ms_b = max_over_ranks(ms_b)
        inv_ms = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in inv_ev])))
        totals = diag().clone().cpu().numpy().tolist()
        again = diag().clone().cpu().numpy().tolist()
        inventory = {"groups": len(w.groups), "tracers_read": len(diag.names),
                     "collective": f"all_reduce(SUM) of {len(w.groups)} doubles over {'NCCL' if world > 1 else 'no other rank (N = 1)'}",
                     "every": "stage", "ms_per_stage_inventory_and_allreduce": inv_ms,
                     "ms_per_step_with_inventory": ms_b / args.steps,
                     "value_with_inventory": global_cells * args.steps / (ms_b * 1e-3) / 1e9,
                     "algorithmic_bytes_per_cell": 8 * len(diag.names),
                     "GBs_per_gpu": 8 * len(diag.names) * w.cells / (inv_ms * 1e-3) / 1e9,
                     "totals": totals, "run_to_run_identical": totals == again,
                     "check": inventory_check(w, diag, rank, world, rows_info)}

    # ---- end-to-end through the public API with host buffers ------------------------------------------
    e2e = None
    if not args.no_e2e:
        # pinned host copies of every tracer and tendency: shrink the e2e grid (fewer y rows, same Nx, Nz) when the
        # node's RAM cannot hold them for every local rank — the leg is PCIe-bound, so Gcell/s is size-independent
        import psutil
        we, note = w, None
        if w.kind != "carbon":
            need = (len(w.model.tracers) + w.nG) * w.model.tracers["P"].data.numel() * 8
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
            budget = 0.5 * psutil.virtual_memory().available / max(1, local_world)
            if need > budget:
                ny_e = max(1, int(w.grid.Ny * budget / need))
                we = Workload(name, device, 1.0, rows=(0, ny_e, ny_e))
                note = f"host RAM limits pinned buffers: e2e grid reduced to {we.grid.Nx}x{we.grid.Ny}x{we.grid.Nz} per GPU"
        hs = HostStage(we, args.copy_engine, args.e2e_slabs)
        for _ in range(2):
            hs.step()
        barrier()
        k = max(2, min(args.steps, 5))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            hs.step()
        b.record()
        barrier()
        my_ms = a.elapsed_time(b)
        ems = max_over_ranks(my_ms)
        e2e_cells = we.cells * world
        equal = None
        # Strong scaling over ranks whose host links are not equally fast (the 8-GPU box: two groups of four GPUs, 1.43×
        # apart, profiles/r04_e2e_probe_n8.jsonl): a stage ends with the slowest link.  Re-share the rows of the ONE grid in
        # proportion to each rank's rate just measured (oceanbiome_b200.distributed.slab_ranges_by_rate) and time again;
        # the device-resident `value` keeps equal slabs.  Both numbers are reported.
        if world > 1 and rows_info is not None and we is w and w.kind != "carbon" and not args.no_e2e_balance:
            from oceanbiome_b200.distributed import slab_ranges_by_rate
            t = torch.tensor([my_ms], dtype=torch.float64, device=device)
            allms = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allms, t)
            times = [x.item() for x in allms]
            nyg = rows_info[2]
            ranges = slab_ranges_by_rate(nyg, [1.0 / x for x in times], minimum=max(1, min(8, nyg // (2 * world))))
            if max(times) > args.e2e_balance_threshold * min(times):
                equal = {"value": e2e_cells * k / (ems * 1e-3) / 1e9, "ms_per_step_per_rank": [round(x / k, 2) for x in times],
                         "rows_per_rank": nyg // world}
                del hs
                j0b, j1b = ranges[rank]
                we = Workload(name, device, args.scale, rows=(j0b, j1b - j0b, nyg))
                hs = HostStage(we, args.copy_engine, args.e2e_slabs)
                for _ in range(2):
                    hs.step()
                barrier()
                a.record()
                for _ in range(k):
                    hs.step()
                b.record()
                barrier()
                my_ms = a.elapsed_time(b)
                ems = max_over_ranks(my_ms)
                e2e_cells = w.grid.Nx * nyg * w.grid.Nz
        e2e = {"value": e2e_cells * k / (ems * 1e-3) / 1e9, "unit": "Gcell-updates/s",
               "h2d_bytes_per_step": hs.h2d_bytes, "d2h_bytes_per_step": hs.d2h_bytes, "steps": k,
               "cells_per_gpu": we.cells, "copy_engine": args.copy_engine, "numa_bound": bool(hs.numa_bound),
               "pcie_GBs_each_direction": max(hs.h2d_bytes, hs.d2h_bytes) * k / (ems * 1e-3) / 1e9}
        if equal is not None:
            e2e["slabs"] = {"how": "rows of the one grid shared out in proportion to each rank's host-link rate measured with equal slabs "
                                   "in this run (a stage ends with the slowest link); bytes and cells_per_gpu are rank 0's",
                            "rows_per_rank": [j1 - j0 for j0, j1 in ranges], "equal_slabs": equal}
        # the leg's ceiling on THIS box, measured with the leg's own pinned buffers and copy pattern on every rank at once:
        # the same slab copies without the kernels between them, then each direction alone (profiles/r04_e2e_probe_*:
        # with 8 ranks the two directions do not overlap on the host side — copies-only ≈ H2D-only + D2H-only — and the
        # ranks behind one host bridge share what it delivers, whatever the copy pattern, slab count or copy engine)
        if we.kind != "carbon":
            def copies(**kw):
                hs.stage.step(kernels=False, **kw)
                barrier()
                a.record()
                for _ in range(2):
                    hs.stage.step(kernels=False, **kw)
                b.record()
                barrier()
                return max_over_ranks(a.elapsed_time(b)) / 2
            c_both, c_in, c_out = copies(), copies(d2h=False), copies(h2d=False)
            e2e["ceiling"] = {
                "how": "the stage's own slab copies of its own pinned buffers with no kernels between them, all ranks at once "
                       "(max over ranks), and each direction alone",
                "copies_only_ms": c_both, "h2d_only_ms": c_in, "d2h_only_ms": c_out, "stage_ms": ems / k,
                "value": we.cells * world / (c_both * 1e-3) / 1e9, "unit": "Gcell-updates/s",
                "frac": c_both / (ems / k),
                "directions_overlap": (c_in + c_out) / c_both,  # 2 = full duplex, 1 = the directions serialise
                "GBs_h2d_alone_per_rank": hs.h2d_bytes / (c_in * 1e-3) / 1e9,
                "GBs_d2h_alone_per_rank": hs.d2h_bytes / (c_out * 1e-3) / 1e9}
        # and the link's own peak: EVERY rank copies one large buffer in both directions at the same time
        del hs
        barrier()
        peak_c = pcie_peak_gbs(device, reps=8 if world > 1 else 3, barrier=barrier if world > 1 else None)
        hostbw_c = host_memcpy_gbs()
        barrier()
        if world > 1:
            t = torch.tensor([peak_c, hostbw_c], dtype=torch.float64, device=device)
            allp = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allp, t)
            peaks = [x[0].item() for x in allp]
            hostbws = [x[1].item() for x in allp]
        else:
            peaks, hostbws = [peak_c], [hostbw_c]
        if world > 1:  # and with the other ranks idle, for the contrast
            if rank == 0:
                solo, host_solo = pcie_peak_gbs(device), host_memcpy_gbs()
            barrier()
        else:
            solo, host_solo = peak_c, hostbw_c
        if rank == 0:
            e2e["pcie_peak_GBs_each_direction"] = float(np.mean(peaks))
            e2e["pcie_peak_per_rank"] = [round(p, 2) for p in peaks]
            e2e["pcie_peak_how"] = (f"1 GiB pinned copies (H2D and D2H) in flight on each of the {world} ranks AT THE SAME TIME "
                                    "(buffers pinned first, then a barrier, then 8 copies per direction back to back); mean over ranks")
            e2e["pcie_frac"] = e2e["pcie_GBs_each_direction"] / e2e["pcie_peak_GBs_each_direction"]
            e2e["pcie_peak_solo_GBs_each_direction"] = solo
            e2e["host_memcpy_GBs_per_rank_concurrent"] = float(np.mean(hostbws))
            e2e["host_memcpy_GBs_solo"] = host_solo
            e2e["limiter"] = ("PCIe link of the one GPU" if world == 1 else
                              f"shared host side: {world} ranks copying at once get {np.mean(peaks):.1f} GB/s per direction each "
                              f"({world * np.mean(peaks):.0f} GB/s per direction in aggregate) against {solo:.1f} GB/s for one rank alone")
            c = e2e.get("ceiling")
            if c and world > 1:
                e2e["limiter"] += (f"; with the stage's own buffers the slowest rank moves {c['GBs_h2d_alone_per_rank']:.1f} GB/s in alone, "
                                   f"{c['GBs_d2h_alone_per_rank']:.1f} GB/s out alone, and both together take {c['copies_only_ms']:.0f} ms against "
                                   f"{c['h2d_only_ms']:.0f} + {c['d2h_only_ms']:.0f} ms (overlap factor {c['directions_overlap']:.2f} of 2): "
                                   f"the stage runs at {c['frac']:.2f} of that copies-only ceiling")
        if note:
            e2e["note"] = note
        we = None  # (it may be `w` itself: drop the reference so that the weak leg below can free the slab)

    # ---- weak-scaling sub-record (N > 1): one full-size grid per GPU ---------------------------------------------------
    weak = None
    if world > 1 and mode == "strong" and not args.no_weak:
        cells_strong = w.cells
        del w
        if inventory is not None:
            del diag
        torch.cuda.empty_cache()
        ww = Workload(name, device, args.scale)
        for _ in range(args.warmup):
            ww.step()
        barrier()
        kw_steps = max(3, min(args.steps, 10))
        wms, _ = timed_steps(ww, kw_steps, False, False, None)
        wms = max_over_ranks(wms)
        weak = {"scaling": "weak", "cells_per_gpu": ww.cells, "steps": kw_steps, "ms_per_step": wms / kw_steps,
                "value": ww.cells * world * kw_steps / (wms * 1e-3) / 1e9, "unit": "Gcell-updates/s"}
        w_cells, w_obj = cells_strong, ww
    else:
        w_cells, w_obj = w.cells, w

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    alg_bytes = w_obj.tendency_bytes_per_cell * w_cells
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(name, w_cells)
    step_ms = ms / args.steps
    kernels = []
    for label, (kname, nbytes, bound) in w_obj.stage_kernels.items():
        if label in ktimes:
            gbs = nbytes * w_cells / (ktimes[label] * 1e-3) / 1e9
            kernels.append({"kernel": kname, "hook": label, "algorithmic_bytes_per_cell": nbytes, "ms": ktimes[label],
                            "achieved": gbs, "unit": "GB/s", "frac": gbs / peak, "bound": bound,
                            "share_of_step": None if use_graph else ktimes[label] / step_ms})
    stage_gbs = w_obj.stage_bytes_per_cell * w_cells / (step_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": w_obj.stage_kernels["tendencies"][0], "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_cell": w_obj.tendency_bytes_per_cell, "kernel_ms": kernel_ms,
                # graph mode: kernel_ms comes from separate eager launches (it includes their launch latency, which the
                # graphed stage does not pay), so a share of the graphed step would be meaningless
                "kernel_share_of_step": None if use_graph else kernel_ms / step_ms,
                "fp64": fp64_roofline(name, w_cells, kernel_ms),
                # the whole stage (every launch of the step, gaps included) against the same HBM peak
                "stage": {"algorithmic_bytes_per_cell": w_obj.stage_bytes_per_cell, "ms": step_ms, "achieved": stage_gbs,
                          "unit": "GB/s", "frac": stage_gbs / peak,
                          "how": "sum of the kernels' algorithmic bytes x cells / ms_per_step (per GPU; max over ranks)"},
                "kernels": kernels}
    cpu = cpu_fused = None
    if not args.no_cpu_baseline and world == 1:
        # ≈ 10 s of one host core on the GPU box: PISCES ≈ 0.13, LOBSTER ≈ 2, the carbonate solve ≈ 0.4 Mcell/s
        budget = {"pisces": 1_500_000, "carbon": 4_000_000}.get(cfg["model"], 8_000_000)
        v, sample, _ = cpu_sample(name, 1, budget)
        cpu = {"value": v / 1e9, "unit": "Gcell-updates/s", "cores": 1, "kind": "port", "sample": sample}
        cpu_fused = fused_cpu_sample(name, 1, budget)
    line = {
        "metric": "BGC tendency Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": mode if world > 1 else args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(name, world, args.scaling),
        "cells_per_gpu": w_cells, "cuda_graph": bool(use_graph),
        "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
        "cpu_baseline_fused": cpu_fused, "inventory": inventory, "weak": weak,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
