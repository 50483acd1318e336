"""pyoracle — numpy/ctypes front-end of the CPU ORACLE (oracle/src/*.c → oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs as the checker or as the timed CPU baseline — never by the product package.
All arrays are float64 numpy parent arrays in the same halo'd layout the CUDA library uses
(shape (Nz+2Hz, Ny+2Hy, Nx+2Hx), x fastest); the struct definitions are those of include/obm_b200.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from oceanbiome_b200 import _lib as abi  # noqa: E402  (struct layouts only — no compute)

LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None
dp = C.POINTER(C.c_double)


def build(force: bool = False):
    """Compile the oracle with its Makefile (gcc -O2 -ffp-contract=off -fopenmp)."""
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, "src", f)) > os.path.getmtime(LIB_PATH)
            for f in os.listdir(os.path.join(_HERE, "src"))):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_carbon_chemistry.restype = C.c_double
        _lib.orc_carbon_chemistry.argtypes = [C.c_double] * 4 + [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double,
                                                                 C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int),
                                                                 C.POINTER(C.c_int)]
        for name in ("orc_K0",):
            getattr(_lib, name).restype = C.c_double
            getattr(_lib, name).argtypes = [C.c_double, C.c_double]
        for name in ("orc_K1", "orc_K2", "orc_KB", "orc_KW", "orc_KP1", "orc_KP2", "orc_KP3", "orc_KSP_calcite",
                     "orc_KSP_aragonite"):
            getattr(_lib, name).restype = C.c_double
            getattr(_lib, name).argtypes = [C.c_double, C.c_double, C.c_int, C.c_double]
        _lib.orc_KS.restype = C.c_double
        _lib.orc_KS.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
        _lib.orc_KF.restype = C.c_double
        _lib.orc_KF.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
        _lib.orc_KSi.restype = C.c_double
        _lib.orc_KSi.argtypes = [C.c_double, C.c_double, C.c_double]
        _lib.orc_ionic_strength.restype = C.c_double
        _lib.orc_ionic_strength.argtypes = [C.c_double]
        _lib.orc_pressure_correction.restype = C.c_double
        _lib.orc_pressure_correction.argtypes = [C.c_int, C.c_double, C.c_double]
        _lib.orc_first_virial.restype = C.c_double
        _lib.orc_first_virial.argtypes = [C.c_double]
        _lib.orc_cross_virial.restype = C.c_double
        _lib.orc_cross_virial.argtypes = [C.c_double]
        _lib.orc_teos10_polynomial_approximation.restype = C.c_double
        _lib.orc_teos10_polynomial_approximation.argtypes = [C.c_double] * 3
        _lib.orc_numerical_mean.restype = C.c_double
        _lib.orc_numerical_mean.argtypes = [dp, dp, C.c_int, C.c_double, C.c_double]
    return _lib


def set_threads(n: int):
    """OpenMP thread count of the oracle (cpu_baseline.cores)."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


class Grid:
    """Host twin of oceanbiome_b200.RectilinearGrid for the oracle: same sizes/halos, host z arrays."""

    def __init__(self, Nx, Ny, Nz, Hx, Hy, Hz, zc_parent, zf_parent, bottom_indices=None):
        self.Nx, self.Ny, self.Nz, self.Hx, self.Hy, self.Hz = Nx, Ny, Nz, Hx, Hy, Hz
        self.zc_parent = np.ascontiguousarray(zc_parent, dtype=np.float64)
        self.zf_parent = np.ascontiguousarray(zf_parent, dtype=np.float64)
        # immersed boundary: int64 x–y parent plane of 1-based bottom-most active cells (or None)
        self.bottom_indices = None if bottom_indices is None else np.ascontiguousarray(bottom_indices, dtype=np.int64)

    @classmethod
    def like(cls, g):
        b = getattr(g, "bottom_indices", None)
        return cls(g.Nx, g.Ny, g.Nz, g.Hx, g.Hy, g.Hz, g.zc_host, g.zf_host, None if b is None else b.cpu().numpy())

    @property
    def parent_shape(self):
        return (self.Nz + 2 * self.Hz, self.Ny + 2 * self.Hy, self.Nx + 2 * self.Hx)

    @property
    def plane_shape(self):
        return (1, self.Ny + 2 * self.Hy, self.Nx + 2 * self.Hx)

    def interior(self, a):
        if a.shape[0] == 1:
            return a[:, self.Hy:self.Hy + self.Ny, self.Hx:self.Hx + self.Nx]
        return a[self.Hz:self.Hz + self.Nz, self.Hy:self.Hy + self.Ny, self.Hx:self.Hx + self.Nx]

    def c_grid(self, i0=0, i1=0, j0=0, j1=0):
        return abi.obm_grid(self.Nx, self.Ny, self.Nz, self.Hx, self.Hy, self.Hz, i0, i1, j0, j1,
                            self.zc_parent.ctypes.data, self.zf_parent.ctypes.data,
                            None if self.bottom_indices is None else self.bottom_indices.ctypes.data)


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _table(arrays):
    return (C.c_void_p * len(arrays))(*[C.c_void_p(_ptr(a)) if a is not None else None for a in arrays])


def _check(arrs):
    for a in arrs:
        if a is not None:
            assert a.dtype == np.float64 and a.flags.c_contiguous, "oracle needs contiguous float64 parent arrays"


# ---- NPD --------------------------------------------------------------------------------------------
def npd_tracer_names(params):
    names = ((C.c_char * 16) * abi.OBM_NPD_MAX_TRACERS)()
    n = lib().orc_npd_layout(C.byref(params), None, names)
    assert n >= 0, f"orc_npd_layout → {n}"
    return tuple(bytes(names[i]).split(b"\0")[0].decode("utf-8") for i in range(n))


def npd_tendencies(grid: Grid, params, tracers, PAR, G=None, accumulate=False):
    """One full-grid pass per tracer (reference launch structure).  Returns list of G parent arrays."""
    _check(list(tracers) + [PAR])
    if G is None:
        G = [np.zeros(grid.parent_shape) for _ in tracers]
    cg = grid.c_grid()
    rc = lib().orc_npd_tendencies(C.byref(cg), C.byref(params), _table(tracers), C.c_void_p(_ptr(PAR)), _table(G),
                                  1 if accumulate else 0)
    assert rc == 0, f"orc_npd_tendencies → {rc}"
    return G


def npd_tendency_scales(grid: Grid, params, tracers, PAR):
    """Σ|additive terms| of every tendency — the S of the parity metric |a − b| ≤ 1e-12·max(|b|, S)."""
    _check(list(tracers) + [PAR])
    S = [np.zeros(grid.parent_shape) for _ in tracers]
    cg = grid.c_grid()
    rc = lib().orc_npd_tendency_scales(C.byref(cg), C.byref(params), _table(tracers), C.c_void_p(_ptr(PAR)), _table(S))
    assert rc == 0, f"orc_npd_tendency_scales → {rc}"
    return S


# ---- light ------------------------------------------------------------------------------------------
def par_twoband(grid: Grid, params, P, surface_PAR, PAR=None):
    PAR = np.zeros(grid.parent_shape) if PAR is None else PAR
    s_arr = surface_PAR if isinstance(surface_PAR, np.ndarray) else None
    _check([P, PAR, s_arr])
    cg = grid.c_grid()
    rc = lib().orc_par_twoband(C.byref(cg), C.byref(params), C.c_void_p(_ptr(P)), C.c_void_p(_ptr(s_arr)),
                               C.c_double(0.0 if s_arr is not None else float(surface_PAR)), C.c_void_p(_ptr(PAR)))
    assert rc == 0
    return PAR


def par_multiband(grid: Grid, params, chl_a, chl_b, chl_scale, surface_PAR, bands=None, total=None):
    nb = params.nbands
    bands = [np.zeros(grid.parent_shape) for _ in range(nb)] if bands is None else bands
    total = np.zeros(grid.parent_shape) if total is None else total
    s_arr = surface_PAR if isinstance(surface_PAR, np.ndarray) else None
    _check([chl_a, chl_b, s_arr, total] + bands)
    cg = grid.c_grid()
    rc = lib().orc_par_multiband(C.byref(cg), C.byref(params), C.c_void_p(_ptr(chl_a)), C.c_void_p(_ptr(chl_b)),
                                 C.c_double(chl_scale), C.c_void_p(_ptr(s_arr)),
                                 C.c_double(0.0 if s_arr is not None else float(surface_PAR)), _table(bands),
                                 C.c_void_p(_ptr(total)))
    assert rc == 0
    return bands, total


def euphotic_depth(grid: Grid, PAR, cutoff=1 / 1000):
    zeu = np.zeros(grid.plane_shape)
    _check([PAR])
    cg = grid.c_grid()
    rc = lib().orc_euphotic_depth(C.byref(cg), C.c_void_p(_ptr(PAR)), C.c_double(cutoff), C.c_void_p(_ptr(zeu)))
    assert rc == 0
    return zeu


def mixed_layer_mean(grid: Grid, zmxl, Cfield):
    out = np.zeros(grid.plane_shape)
    _check([zmxl, Cfield])
    cg = grid.c_grid()
    rc = lib().orc_mixed_layer_mean(C.byref(cg), C.c_void_p(_ptr(zmxl)), C.c_void_p(_ptr(Cfield)), C.c_void_p(_ptr(out)))
    assert rc == 0
    return out


def numerical_mean(lam, Cc, lo, hi):
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    Cc = np.ascontiguousarray(Cc, dtype=np.float64)
    return lib().orc_numerical_mean(lam.ctypes.data_as(dp), Cc.ctypes.data_as(dp), len(lam), float(lo), float(hi))


# ---- carbon chemistry -------------------------------------------------------------------------------
def carbon_chemistry(DIC, T, S, Alk=0.0, pH=None, P=None, silicate=0.0, phosphate=0.0, output=abi.CC_FCO2,
                     initial_pH_guess=8.0, return_iters=False):
    ni, nf = C.c_int(), C.c_int()
    v = lib().orc_carbon_chemistry(DIC, T, S, Alk, pH is not None, pH or 0.0, P is not None, P or 0.0, silicate,
                                   phosphate, initial_pH_guess, output, C.byref(ni), C.byref(nf))
    return (v, ni.value, nf.value) if return_iters else v


def carbon_chemistry_sweep(T, S, DIC, Alk, P=None, silicate=None, phosphate=None, pH=None, output=abi.CC_FCO2,
                           initial_pH_guess=8.0):
    arrs = [T, S, DIC, Alk, P, silicate, phosphate, pH]
    _check(arrs)
    out = np.empty_like(DIC)
    fe = C.c_int64()
    rc = lib().orc_carbon_chemistry_sweep(C.c_int64(DIC.size), *[C.c_void_p(_ptr(a)) for a in arrs],
                                          C.c_double(initial_pH_guess), C.c_int(output), C.c_void_p(_ptr(out)),
                                          C.byref(fe))
    assert rc == 0
    return out, fe.value


def calcite_saturation(grid: Grid, T, S, DIC, Alk, Si):
    Om = np.zeros(grid.parent_shape)
    _check([T, S, DIC, Alk, Si])
    cg = grid.c_grid()
    rc = lib().orc_calcite_saturation(C.byref(cg), *[C.c_void_p(_ptr(a)) for a in (T, S, DIC, Alk, Si, Om)])
    assert rc == 0
    return Om


# ---- negative tracers / inventory ----------------------------------------------------------------
def make_groups(names, groups):
    """groups: list of (tracer_names, scalefactors) → ctypes obm_scale_group array indexing `names`."""
    arr = (abi.obm_scale_group * len(groups))()
    for q, (tn, sf) in enumerate(groups):
        arr[q].n = len(tn)
        for m, (t, f) in enumerate(zip(tn, sf)):
            arr[q].index[m] = list(names).index(t)
            arr[q].scalefactor[m] = float(f)
    return arr


def scale_negative_tracers(grid: Grid, tracers, groups, invalid_fill_value=float("nan")):
    """In place on the list of parent arrays; `groups` from make_groups.  One pass per group."""
    _check(tracers)
    cg = grid.c_grid()
    rc = lib().orc_scale_negative_tracers(C.byref(cg), len(tracers), _table(tracers), len(groups), groups,
                                          C.c_double(invalid_fill_value))
    assert rc == 0
    return tracers


def zero_negative_tracers(tracers):
    _check(tracers)
    rc = lib().orc_zero_negative_tracers(C.c_int64(tracers[0].size), len(tracers), _table(tracers))
    assert rc == 0
    return tracers


def inventory(grid: Grid, tracers, groups, cell_volume=None, uniform_volume=1.0):
    _check(list(tracers) + [cell_volume])
    out = np.zeros(len(groups))
    cg = grid.c_grid()
    rc = lib().orc_inventory(C.byref(cg), len(tracers), _table(tracers), len(groups), groups,
                             C.c_void_p(_ptr(cell_volume)), C.c_double(uniform_volume), C.c_void_p(_ptr(out)))
    assert rc == 0
    return out


# ---- PISCES -------------------------------------------------------------------------------------------
def pisces_fields(aux: dict):
    """aux: dict of numpy parent arrays keyed like obm_pisces_fields members."""
    f = abi.obm_pisces_fields()
    for k in ("PAR1", "PAR2", "PAR3", "PAR", "Omega", "wPOC", "wGOC", "mixed_layer_depth_xy", "euphotic_depth_xy",
              "mean_mixed_layer_vertical_diffusivity_xy", "mean_mixed_layer_light_xy"):
        _check([aux[k]])
        setattr(f, k, aux[k].ctypes.data)
    return f


def pisces_tendencies(grid: Grid, params, tracers, aux: dict, G=None, accumulate=False, skip=("T", "S"), rows=None):
    """One full-grid pass per tracer (reference launch structure).  tracers: list of 26 parent arrays.
    `rows = (j0, j1)`: only those interior rows (ModelLatitude: one call per row with that row's parameter block)."""
    _check(tracers)
    if G is None:
        G = [np.zeros(grid.parent_shape) if n < 24 else None for n in range(abi.OBM_PISCES_NTRACERS)]
    f = pisces_fields(aux)
    cg = grid.c_grid() if rows is None else grid.c_grid(j0=rows[0], j1=rows[1])
    rc = lib().orc_pisces_tendencies(C.byref(cg), C.byref(params), _table(tracers), C.byref(f), _table(G),
                                     1 if accumulate else 0)
    assert rc == 0, f"orc_pisces_tendencies → {rc}"
    return G


def pisces_tendency_scales(grid: Grid, params, tracers, aux: dict, S=None, rows=None):
    """Σ|additive terms| of each of the 24 tendencies — the S of the parity metric."""
    _check(tracers)
    if S is None:
        S = [np.zeros(grid.parent_shape) if n < 24 else None for n in range(abi.OBM_PISCES_NTRACERS)]
    f = pisces_fields(aux)
    cg = grid.c_grid() if rows is None else grid.c_grid(j0=rows[0], j1=rows[1])
    rc = lib().orc_pisces_tendency_scales(C.byref(cg), C.byref(params), _table(tracers), C.byref(f), _table(S))
    assert rc == 0, f"orc_pisces_tendency_scales → {rc}"
    return S


def set_nested_free_iron_scale(on: bool):
    """S of SFe, BFe, Fe with Fe′'s OWN Σ|terms| in place of |Fe′| (oracle_pisces.c::free_iron): for the test of extreme
    finite states, where Fe′ = −Δ + √(Δ² + 4K·Fe) is cancellation noise.  Off by default; switch it back off after use."""
    lib().orc_pisces_set_nested_free_iron_scale(1 if on else 0)


_select_lib = None


def select_lib():
    """The same oracle built with -DORC_SELECT_MINMAX: Julia's NaN-propagating min / max replaced by compare + select,
    i.e. the arithmetic of the kernels' fast pass (used to show where that pass could swallow a NaN)."""
    global _select_lib
    if _select_lib is None:
        path = os.path.join(_HERE, "_build", "liboracle_select.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(LIB_PATH):
            subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)
        _select_lib = C.CDLL(path)
    return _select_lib


def pisces_point(params, values, PAR1, PAR2, PAR3, PAR, Omega, wPOC, wGOC, zmxl, zeu, kappa, mlPAR, z, select=False):
    vals = (C.c_double * abi.OBM_PISCES_NTRACERS)(*values)
    out = (C.c_double * abi.OBM_PISCES_NTRACERS)()
    fn = (select_lib() if select else lib()).orc_pisces_point
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 12 + [C.c_void_p]
    fn(C.byref(params), vals, PAR1, PAR2, PAR3, PAR, Omega, wPOC, wGOC, zmxl, zeu, kappa, mlPAR, z, out)
    return list(out)


def pisces_point_terms(params, values, PAR1, PAR2, PAR3, PAR, Omega, wPOC, wGOC, zmxl, zeu, kappa, mlPAR, z):
    """(tendencies, Σ|terms| per tendency) of one point."""
    vals = (C.c_double * abi.OBM_PISCES_NTRACERS)(*values)
    out = (C.c_double * abi.OBM_PISCES_NTRACERS)()
    sc = (C.c_double * abi.OBM_PISCES_NTRACERS)()
    fn = lib().orc_pisces_point_terms
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 12 + [C.c_void_p, C.c_void_p]
    fn(C.byref(params), vals, PAR1, PAR2, PAR3, PAR, Omega, wPOC, wGOC, zmxl, zeu, kappa, mlPAR, z, out, sc)
    return list(out), list(sc)


def cbm_day_length(t, phi):
    fn = lib().orc_cbm_day_length
    fn.restype = C.c_double
    fn.argtypes = [C.c_double, C.c_double]
    return fn(t, phi)


# ---- sediments ----------------------------------------------------------------------------------------
def sediment_fields(**ptrs):
    """Build an obm_sediment_fields from numpy arrays: NO3, NH4, O2, sinking=[…], sinking_w=[…], pools=[…], Gn=[…],
    Gm=[…], tracked=[…], G_coupled=[…], bottom_indices=int64 array."""
    f = abi.obm_sediment_fields()
    keep = []
    for k in ("NO3", "NH4", "O2"):
        if ptrs.get(k) is not None:
            setattr(f, k, ptrs[k].ctypes.data)
            keep.append(ptrs[k])
    if ptrs.get("bottom_indices") is not None:
        f.bottom_indices_xy = ptrs["bottom_indices"].ctypes.data
    for name, member in (("sinking", "sinking"), ("sinking_w", "sinking_w"), ("pools", "pools"), ("Gn", "Gn"), ("Gm", "Gm"),
                         ("tracked", "tracked_xy"), ("G_coupled", "G_coupled")):
        for n, a in enumerate(ptrs.get(name) or []):
            if a is not None:
                getattr(f, member)[n] = a.ctypes.data
    return f


def sediment_update_state(grid: Grid, params, fields, dt, chi=0.1):
    cg = grid.c_grid()
    rc = lib().orc_sediment_update_state(C.byref(cg), C.byref(params), C.byref(fields), C.c_double(dt), C.c_double(chi))
    assert rc == 0


def sediment_update_tendencies(grid: Grid, params, fields):
    cg = grid.c_grid()
    rc = lib().orc_sediment_update_tendencies(C.byref(cg), C.byref(params), C.byref(fields))
    assert rc == 0


def find_bottom_cells(grid: Grid, bottom_height):
    out = np.ones(grid.plane_shape, dtype=np.int64)
    cg = grid.c_grid()
    rc = lib().orc_find_bottom_cells(C.byref(cg), C.c_void_p(_ptr(bottom_height)), C.c_void_p(out.ctypes.data))
    assert rc == 0
    return out


def sediment_point(params, pools, NO3, NH4, O2, fN, fC=0.0):
    """→ (pool tendencies, coupled fluxes) at one point."""
    L = lib()
    for fn in (L.orc_sediment_pool_tendency, L.orc_sediment_coupled_flux):
        fn.restype = C.c_double
        fn.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 5 + [C.c_int]
    pl = (C.c_double * 6)(*(list(pools) + [0.0] * (6 - len(pools))))
    smg = params.model == abi.SED_SIMPLE_MULTI_G
    npool = (6 if params.carbon else 3) if smg else 1
    nc = (4 if params.carbon else 3) if smg else 1
    return ([L.orc_sediment_pool_tendency(C.byref(params), pl, NO3, NH4, O2, fN, fC, n) for n in range(npool)],
            [L.orc_sediment_coupled_flux(C.byref(params), pl, NO3, NH4, O2, fN, fC, n) for n in range(nc)])


# ---- gas exchange ---------------------------------------------------------------------------------------
def gas_exchange_flux(grid: Grid, params, T, S, tracer=None, DIC=None, Alk=None, silicate=None, phosphate=None,
                      wind_speed=None, air_concentration=None, G_top=None):
    """→ flux plane (grid.plane_shape); G_top (3-D parent) is updated in place when given."""
    arrs = [T, S, tracer, DIC, Alk, silicate, phosphate, wind_speed, air_concentration]
    _check(arrs + [G_top])
    flux = np.zeros(grid.plane_shape)
    cg = grid.c_grid()
    rc = lib().orc_gas_exchange_flux(C.byref(cg), C.byref(params), *[C.c_void_p(_ptr(a)) for a in arrs],
                                     C.c_void_p(_ptr(flux)), C.c_void_p(_ptr(G_top)))
    assert rc == 0
    return flux


def gas_exchange_point(params, T, S, tracer=0.0, DIC=0.0, Alk=0.0, silicate=0.0, phosphate=0.0, u10=None, air=None):
    fn = lib().orc_gas_exchange_point
    fn.restype = C.c_double
    fn.argtypes = [C.c_void_p] + [C.c_double] * 9
    return fn(C.byref(params), T, S, tracer, DIC, Alk, silicate, phosphate,
              params.wind_speed if u10 is None else u10, params.air_concentration if air is None else air)


def polynomial(coefficients, x):
    fn = lib().orc_polynomial
    fn.restype = C.c_double
    fn.argtypes = [C.c_int, dp, C.c_double]
    c = (C.c_double * len(coefficients))(*coefficients)
    return fn(len(coefficients) - 1, c, x)


def transfer_velocity(params, u10, T, S):
    fn = lib().orc_transfer_velocity
    fn.restype = C.c_double
    fn.argtypes = [C.c_void_p] + [C.c_double] * 3
    return fn(C.byref(params), u10, T, S)


def w92_solubility(coefficients, T, S):
    fn = lib().orc_w92_solubility
    fn.restype = C.c_double
    fn.argtypes = [dp, C.c_double, C.c_double]
    return fn((C.c_double * 6)(*coefficients), T, S)


# ---- time stepping ----------------------------------------------------------------------------------------
def rk3_substep(grid: Grid, U, Gn, Gm, dt, gamma, zeta=None, cache_previous=True):
    """In place on the lists of parent arrays U, Gm."""
    _check(list(U) + list(Gn) + list(Gm))
    cg = grid.c_grid()
    rc = lib().orc_rk3_substep(C.byref(cg), len(U), _table(U), _table(Gn), _table(Gm), C.c_double(dt), C.c_double(gamma),
                               C.c_double(0.0 if zeta is None else zeta), int(zeta is not None), int(cache_previous))
    assert rc == 0


# ---- sinking ------------------------------------------------------------------------------------------------
def sinking_tendencies(grid: Grid, tracers, w_faces, G, advection=0, accumulate=True):
    """G[t] (+)= −∂z(w c), in place on the list of parent arrays G (oracle_sinking.c)."""
    _check(list(tracers) + list(w_faces) + list(G))
    cg = grid.c_grid()
    rc = lib().orc_sinking_tendencies(C.byref(cg), len(tracers), _table(tracers), _table(w_faces), _table(G),
                                      int(advection), int(accumulate))
    assert rc == 0
    return G


# ---- particles / sugar kelp -------------------------------------------------------------------------------------
KELP_NAMES = ("A", "N", "C", "NO₃", "NH₄", "DIC", "O₂", "DOC", "DON", "bPOC", "bPON")


def _kelp_protos():
    L = lib()
    if getattr(L, "_kelp_ready", False):
        return L
    L.orc_kelp.restype = C.c_double
    L.orc_kelp.argtypes = [C.POINTER(abi.obm_sugar_kelp_params), C.c_int] + [C.c_double] * 11
    L.orc_kelp_seasonal_limitation.restype = C.c_double
    L.orc_kelp_seasonal_limitation.argtypes = [C.POINTER(abi.obm_sugar_kelp_params), C.c_double]
    L.orc_kelp_light_inhibition.restype = C.c_double
    L.orc_kelp_light_inhibition.argtypes = [C.POINTER(abi.obm_sugar_kelp_params), C.c_double, C.POINTER(C.c_int)]
    L.orc_particle_cell.restype = C.c_int64
    L.orc_particle_cell.argtypes = [C.POINTER(abi.obm_grid), C.POINTER(abi.obm_particles), C.c_int64]
    L.orc_kelp_update_tendencies.restype = C.c_int
    L.orc_kelp_update_tendencies.argtypes = [C.POINTER(abi.obm_grid), C.POINTER(abi.obm_sugar_kelp_params),
                                             C.POINTER(abi.obm_particles), C.POINTER(abi.obm_kelp_tracers), C.c_void_p, C.c_double]
    L.orc_kelp_step.restype = C.c_int
    L.orc_kelp_step.argtypes = [C.POINTER(abi.obm_grid), C.POINTER(abi.obm_sugar_kelp_params), C.POINTER(abi.obm_particles),
                                C.POINTER(abi.obm_kelp_tracers), C.c_double, C.c_double, C.c_void_p]
    L._kelp_ready = True
    return L


def kelp(params, name, t, A, N, Cr, T, NO3, NH4, PAR, u=0.0, v=0.0, w=0.0):
    """`kelp(Val(name), t, A, N, C, u, v, w, T, NO₃, NH₄, PAR)` for name in KELP_NAMES (equations.jl, coupling.jl)."""
    return _kelp_protos().orc_kelp(C.byref(params), KELP_NAMES.index(name), t, A, N, Cr, u, v, w, T, NO3, NH4, PAR)


def kelp_seasonal_limitation(params, t):
    return _kelp_protos().orc_kelp_seasonal_limitation(C.byref(params), t)


def kelp_light_inhibition(params, Pm):
    """β of solve_for_light_inhibition with the reference's solver; returns (β, iterations taken)."""
    n = C.c_int(0)
    b = _kelp_protos().orc_kelp_light_inhibition(C.byref(params), Pm, C.byref(n))
    return b, n.value


def make_particles(x, y, z, A, N, Cr, scalefactors, x0, dx, y0, dy, topology):
    """obm_particles over host arrays (kept alive by the returned tuple)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (x, y, z, A, N, Cr, scalefactors)]
    q = abi.obm_particles()
    q.n = len(arrs[0])
    for name, a in zip(("x", "y", "z", "A", "N", "C", "scalefactors"), arrs):
        setattr(q, name, a.ctypes.data if a is not None else None)
    q.x0, q.dx, q.y0, q.dy = x0, dx, y0, dy
    for c in range(3):
        q.topology[c] = topology[c]
    return q, arrs


def make_kelp_tracers(T, NO3, NH4, PAR, u=None, v=None, w=None):
    f = abi.obm_kelp_tracers()
    keep = []
    for name, a in (("T", T), ("NO3", NO3), ("NH4", NH4), ("PAR", PAR), ("u", u), ("v", v), ("w", w)):
        if a is not None:
            _check([a])
            keep.append(a)
            setattr(f, name, a.ctypes.data)
    return f, keep


def particle_cell(grid: Grid, q, n):
    cg = grid.c_grid()
    return _kelp_protos().orc_particle_cell(C.byref(cg), C.byref(q), n)


def kelp_update_tendencies(grid: Grid, params, q, f, G, t):
    """G: list of 8 parent arrays (or None) in coupled_tracers order; updated in place."""
    _check([g for g in G if g is not None])
    cg = grid.c_grid()
    rc = _kelp_protos().orc_kelp_update_tendencies(C.byref(cg), C.byref(params), C.byref(q), C.byref(f), _table(G), t)
    assert rc == 0


def kelp_step(grid: Grid, params, q, f, t, dt, tendencies_out=None):
    cg = grid.c_grid()
    rc = _kelp_protos().orc_kelp_step(C.byref(cg), C.byref(params), C.byref(q), C.byref(f), t, dt,
                                      _table(tendencies_out) if tendencies_out is not None else None)
    assert rc == 0
