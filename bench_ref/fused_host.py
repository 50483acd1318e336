"""fused_host — ctypes front-end of bench_ref/_build/libobm_fused_host.so: the GPU kernels' fused algorithm compiled for
the host from the kernels' own headers (bench_ref/fused_host.cpp).

MEASUREMENT AND TEST INFRASTRUCTURE ONLY: used by bench.py's `cpu_baseline_fused` leg (the second CPU column,
which separates the algorithmic gain — fusion, fixed-iteration Newton — from the hardware gain) and by
tests/test_fused_host.py (the kernel source's operation order against the oracle, without a GPU).  The product package
neither imports nor links it.  Arrays are float64 numpy parent arrays in the halo'd layout of include/obm_b200.h."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
from oceanbiome_b200 import _lib as abi  # noqa: E402  (struct layouts only)

LIB_PATH = os.path.join(_HERE, "_build", "libobm_fused_host.so")
_lib = None


def build(force: bool = False):
    srcs = [os.path.join(_HERE, "fused_host.cpp")] + [os.path.join(ROOT, "oceanbiome.jl_b200", "csrc", f)
                                                      for f in ("pisces_cell.cuh", "carbon_chemistry.cuh", "obm_common.cuh")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _table(arrays):
    return (C.c_void_p * len(arrays))(*[C.c_void_p(_ptr(a)) if a is not None else None for a in arrays])


def pisces_tendencies(grid, params, tracers, aux: dict, G=None, accumulate=False, exact=False):
    """The kernel's `cell_tendencies<EXACT>` over a host grid (`grid`: pyoracle.Grid).  Returns the 26 G parent arrays."""
    import pyoracle
    if G is None:
        G = [np.zeros(grid.parent_shape) if n < 24 else None for n in range(abi.OBM_PISCES_NTRACERS)]
    f = pyoracle.pisces_fields(aux)
    cg = grid.c_grid()
    rc = lib().fused_pisces_tendencies(C.byref(cg), C.byref(params), _table(tracers), C.byref(f), _table(G),
                                       1 if accumulate else 0, 1 if exact else 0)
    assert rc == 0, f"fused_pisces_tendencies → {rc}"
    return G


def scale_negative_tracers_calcite_saturation(grid, tracers, groups, T, S, DIC, Alk, Si, Omega=None, fill=float("nan"),
                                              level_tables=True):
    Omega = np.zeros(grid.parent_shape) if Omega is None else Omega
    cg = grid.c_grid()
    rc = lib().fused_scale_negative_tracers_calcite_saturation(
        C.byref(cg), len(tracers), _table(tracers), len(groups), groups, C.c_double(fill), C.c_void_p(_ptr(T)), C.c_void_p(_ptr(S)),
        C.c_void_p(_ptr(DIC)), C.c_void_p(_ptr(Alk)), C.c_void_p(_ptr(Si)), C.c_void_p(_ptr(Omega)), 1 if level_tables else 0)
    assert rc == 0
    return Omega


def par_multiband_column_state(grid, params, chl_a, chl_b, chl_scale, surface, zmxl, cutoff=1 / 1000):
    nb = params.nbands
    bands = [np.zeros(grid.parent_shape) for _ in range(nb)]
    total = np.zeros(grid.parent_shape)
    zeu, mean = np.zeros(grid.plane_shape), np.zeros(grid.plane_shape)
    cg = grid.c_grid()
    rc = lib().fused_par_multiband_column_state(
        C.byref(cg), C.byref(params), C.c_void_p(_ptr(chl_a)), C.c_void_p(_ptr(chl_b)), C.c_double(chl_scale), C.c_double(surface),
        _table(bands), C.c_void_p(_ptr(total)), C.c_void_p(_ptr(zmxl)), C.c_double(cutoff), C.c_void_p(_ptr(zeu)), C.c_void_p(_ptr(mean)))
    assert rc == 0
    return bands, total, zeu, mean


def carbon_chemistry_sweep(T, S, DIC, Alk, output=abi.CC_PH_FREE, iterations=12):
    out = np.empty_like(T)
    rc = lib().fused_carbon_chemistry(C.c_longlong(T.size), C.c_void_p(_ptr(T)), C.c_void_p(_ptr(S)), C.c_void_p(_ptr(DIC)),
                                      C.c_void_p(_ptr(Alk)), int(output), int(iterations), C.c_void_p(_ptr(out)))
    assert rc == 0
    return out


def sample(name, threads, budget_cells, cfg):
    """bench.py's `cpu_baseline_fused`: one stage of the workload on a bounded sub-volume with the FUSED algorithm on
    `threads` host threads → dict for the JSON line, or None for workloads whose kernel core is not host-buildable."""
    import pyoracle
    import oceanbiome_b200 as ob
    from oceanbiome_b200 import synthetic
    build()
    pyoracle.set_threads(threads)
    if cfg["model"] == "carbon":
        n = int(budget_cells)
        u = lambda nm: synthetic.uniform_numpy(synthetic.field_id(nm), 0, n)  # noqa: E731
        T, S, DIC = -2.0 + 37.0 * u("T"), 20.0 + 20.0 * u("S"), 1800.0 + 600.0 * u("DIC")
        lo = np.maximum(DIC * 1.02, 2000.0)
        Alk = lo + (2600.0 - lo) * u("Alk")
        t0 = time.perf_counter()
        carbon_chemistry_sweep(T, S, DIC, Alk)
        dt = time.perf_counter() - t0
        return {"value": n / dt / 1e9, "unit": "Gcell-updates/s", "cores": threads, "kind": "fused-algorithm port",
                "sample": f"first {n} cells of the sweep; the kernel's solve (analytic start, <= 12 Newton steps in ln[H+]) compiled for the host"}
    if cfg["model"] != "pisces":
        return None
    import torch
    from oceanbiome_b200.pisces import PISCES, TRACERS, DepthDependantSinkingSpeed, synthetic_range
    Nx, Ny, Nz = cfg["size"]
    ny = max(1, min(Ny, int(budget_cells // (Nx * Nz))))
    Lx, Ly, Lz = cfg["extent"]
    grid = ob.RectilinearGrid(size=(Nx, ny, Nz), x=(0.0, Lx), y=(0.0, Ly * ny / Ny), z=(-Lz, 0.0), device="cpu")
    og = pyoracle.Grid.like(grid)
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
    G = [np.zeros(og.parent_shape) if n < 24 else None for n in range(26)]
    t0 = time.perf_counter()
    Om = scale_negative_tracers_calcite_saturation(og, [host[n] for n in snames], cgroups, host["T"], host["S"], host["DIC"],
                                                   host["Alk"], host["Si"])
    bands, total, zeu, mean = par_multiband_column_state(og, bgc.light_attenuation.c_params(), host["PChl"], host["DChl"], 1.0, 100.0, zmxl)
    aux = {"PAR1": bands[0], "PAR2": bands[1], "PAR3": bands[2], "PAR": total, "Omega": Om, "wPOC": wPOC, "wGOC": wGOC,
           "mixed_layer_depth_xy": zmxl, "euphotic_depth_xy": zeu, "mean_mixed_layer_vertical_diffusivity_xy": kappa,
           "mean_mixed_layer_light_xy": mean}
    pisces_tendencies(og, u.c_params(0.0), [host[n] for n in TRACERS], aux, G=G, accumulate=True)
    dt = time.perf_counter() - t0
    return {"value": grid.ncells / dt / 1e9, "unit": "Gcell-updates/s", "cores": threads, "kind": "fused-algorithm port",
            "sample": f"{grid.Nx}x{grid.Ny}x{grid.Nz} sub-volume (1/{Nx * Ny * Nz / grid.ncells:.0f} of the grid): the three fused passes of the GPU "
                      "stage (scaling + Omega, 3-band PAR with zeu and mean, 24 tendencies) built for the host from the kernels' own headers"}
