"""Light attenuation models — host-side mirror of src/Light/ (same constructor keywords and
defaults).  The column scans run in csrc/light.cu behind obm_par_twoband / obm_par_multiband.

User-supplied surface-PAR callables cannot cross a C ABI, so — like the Julia glue — this layer
evaluates `getbc(surface_PAR, i, j, grid, clock, fields)` (2band.jl:4) into a scalar or a 2-D device
field each stage and passes it as data.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .grids import CenterField, Field, Field2D, RectilinearGrid, current_stream_ptr, require_cuda

hours = 3600.0


def default_surface_PAR(*args):
    """Light.jl:45-47: `100 max(0, cos(t π / 12hours))`, callable as f(t), f(x, t) or f(x, y, t)."""
    t = args[-1]
    return 100 * max(0.0, math.cos(t * math.pi / (12 * hours)))


def evaluate_surface_PAR(surface_PAR, grid: RectilinearGrid, clock, discrete_form=False, parameters=None, fields=None):
    """→ (field_ptr_or_None, scalar).  Accepts a number, a `Field` (2-D, parent x-y layout), a
    torch/numpy (Ny, Nx) array, a continuous-form f(x, y, t[, parameters]) or — with
    `discrete_form=True` — f(i, j, grid, clock, fields[, parameters]) (test/test_light.jl:107-131)."""
    if isinstance(surface_PAR, (int, float)):
        return None, float(surface_PAR)
    if isinstance(surface_PAR, Field):
        return surface_PAR, 0.0
    if torch.is_tensor(surface_PAR) or isinstance(surface_PAR, np.ndarray):
        return Field2D(grid, "surface_PAR").set(surface_PAR), 0.0
    if not callable(surface_PAR):
        raise TypeError(f"unsupported surface_PAR {type(surface_PAR)}")
    t = clock.time if clock is not None else 0.0
    extra = () if parameters is None else (parameters,)
    if surface_PAR is default_surface_PAR:
        return None, float(default_surface_PAR(t))
    xs = (grid.x[0] if grid.x is not None else 0.0) + (np.arange(grid.Nx) + 0.5) * grid.dx
    ys = (grid.y[0] if grid.y is not None else 0.0) + (np.arange(grid.Ny) + 0.5) * grid.dy
    vals = np.empty((grid.Ny, grid.Nx))
    for j in range(grid.Ny):
        for i in range(grid.Nx):
            if discrete_form:
                vals[j, i] = surface_PAR(i + 1, j + 1, grid, clock, fields, *extra)
            else:
                vals[j, i] = surface_PAR(xs[i], ys[j], t, *extra)
    if np.all(vals == vals.flat[0]):
        return None, float(vals.flat[0])
    return Field2D(grid, "surface_PAR").set(vals), 0.0


class TwoBandPhotosyntheticallyActiveRadiation:
    """`TwoBandPhotosyntheticallyActiveRadiation(; grid, …)` — 2band.jl:106-147 (Karleskind et al. 2011)."""

    def __init__(self, grid: RectilinearGrid, water_red_attenuation=0.225, water_blue_attenuation=0.0232,
                 chlorophyll_red_attenuation=0.037, chlorophyll_blue_attenuation=0.074,
                 chlorophyll_red_exponent=0.629, chlorophyll_blue_exponent=0.674, pigment_ratio=0.7,
                 phytoplankton_chlorophyll_ratio=1.31, surface_PAR=default_surface_PAR, discrete_form=False,
                 parameters=None):
        self.grid = grid
        self.water_red_attenuation = water_red_attenuation
        self.water_blue_attenuation = water_blue_attenuation
        self.chlorophyll_red_attenuation = chlorophyll_red_attenuation
        self.chlorophyll_blue_attenuation = chlorophyll_blue_attenuation
        self.chlorophyll_red_exponent = chlorophyll_red_exponent
        self.chlorophyll_blue_exponent = chlorophyll_blue_exponent
        self.pigment_ratio = pigment_ratio
        self.phytoplankton_chlorophyll_ratio = phytoplankton_chlorophyll_ratio
        self.surface_PAR, self.discrete_form, self.parameters = surface_PAR, discrete_form, parameters
        self.field = CenterField(grid, "PAR")

    def c_params(self) -> _lib.obm_twoband_params:
        return _lib.obm_twoband_params(
            self.water_red_attenuation, self.water_blue_attenuation, self.chlorophyll_red_attenuation,
            self.chlorophyll_blue_attenuation, self.chlorophyll_red_exponent, self.chlorophyll_blue_exponent,
            self.pigment_ratio, self.phytoplankton_chlorophyll_ratio)

    def update_biogeochemical_state(self, model, stream: Optional[int] = None):
        """2band.jl:148-155 — uses tracer P directly (not `chlorophyll(bgc, model)`), as the reference does."""
        P = model.tracers["P"]
        require_cuda(P, self.field)
        sfield, sconst = evaluate_surface_PAR(self.surface_PAR, self.grid, model.clock, self.discrete_form,
                                              self.parameters, model.tracers)
        cg, p = self.grid.c_grid(), self.c_params()
        s = stream if stream is not None else current_stream_ptr(self.grid.device)
        rc = _lib.load().obm_par_twoband(C.byref(cg), C.byref(p), P.ptr, sfield.ptr if sfield else None, sconst,
                                         self.field.ptr, s)
        _lib.check(rc, "obm_par_twoband")

    def biogeochemical_auxiliary_fields(self):
        return {"PAR": self.field}  # 2band.jl:160

    def summary(self):
        return "Two-band light attenuation model (Float64)"


# Morel & Maritorena (2001) tables, 350–700 nm every 5 nm — src/Light/morel_coefficients.jl:1-31
MOREL_λ = np.arange(350, 705, 5, dtype=np.float64)
MOREL_kʷ = np.array([
    0.0271, 0.0238, 0.0216, 0.0188, 0.0177, 0.01595, 0.0151, 0.01376, 0.01271, 0.01208,
    0.01042, 0.0089, 0.00812, 0.00765, 0.00758, 0.00768, 0.0077, 0.00792, 0.00885, 0.0099,
    0.01148, 0.01182, 0.01188, 0.01211, 0.01251, 0.0132, 0.01444, 0.01526, 0.0166, 0.01885,
    0.02188, 0.02701, 0.03385, 0.0409, 0.04214, 0.04287, 0.04454, 0.0463, 0.04846, 0.05212,
    0.05746, 0.06053, 0.0628, 0.06507, 0.07034, 0.07801, 0.09038, 0.11076, 0.13584, 0.16792,
    0.2231, 0.25838, 0.26506, 0.26843, 0.27612, 0.284, 0.29218, 0.30176, 0.31134, 0.32553,
    0.34052, 0.3715, 0.41048, 0.42947, 0.43946, 0.44844, 0.46543, 0.48642, 0.5164, 0.55939, 0.62438])
MOREL_e = np.array([
    0.778, 0.767, 0.756, 0.737, 0.72, 0.7, 0.685, 0.673, 0.67, 0.66,
    0.64358, 0.64776, 0.65175, 0.65555, 0.65917, 0.66259, 0.66583, 0.66889, 0.67175, 0.67443,
    0.67692, 0.67923, 0.68134, 0.68327, 0.68501, 0.68657, 0.68794, 0.68903, 0.68955, 0.68947,
    0.6888, 0.68753, 0.68567, 0.6832, 0.68015, 0.67649, 0.67224, 0.66739, 0.66195, 0.65591,
    0.64927, 0.64204, 0.64, 0.63, 0.623, 0.615, 0.61, 0.614, 0.618, 0.622,
    0.626, 0.63, 0.634, 0.638, 0.642, 0.647, 0.653, 0.658, 0.663, 0.667,
    0.672, 0.677, 0.682, 0.687, 0.695, 0.697, 0.693, 0.665, 0.64, 0.62, 0.6])
MOREL_χ = np.array([
    0.153, 0.149, 0.144, 0.14, 0.136, 0.131, 0.127, 0.123, 0.119, 0.118,
    0.11748, 0.12066, 0.12259, 0.12326, 0.12269, 0.12086, 0.11779, 0.11372, 0.10963, 0.1056,
    0.10165, 0.09776, 0.09393, 0.09018, 0.08649, 0.08287, 0.07932, 0.07584, 0.07242, 0.06907,
    0.06579, 0.06257, 0.05943, 0.05635, 0.05341, 0.05072, 0.04829, 0.04611, 0.04419, 0.04253,
    0.04111, 0.03996, 0.039, 0.0375, 0.036, 0.034, 0.033, 0.0328, 0.0325, 0.033,
    0.034, 0.035, 0.036, 0.0375, 0.0385, 0.04, 0.042, 0.043, 0.044, 0.0445,
    0.045, 0.046, 0.0475, 0.049, 0.0515, 0.052, 0.0505, 0.044, 0.039, 0.034, 0.03])


def numerical_mean(λ, C_, idx1, idx2):
    """multi_band.jl:136-140 (0-based inclusive indices here)."""
    integral = 0.0
    for n in range(idx1 + 1, idx2 + 1):
        integral += (C_[n] + C_[n - 1]) * (λ[n] - λ[n - 1]) / 2
    return integral / (λ[idx2] - λ[idx1])


def par_symbol(n: int) -> str:
    """`par_symbol(n)` multi_band.jl:142 — PAR₁, PAR₂, …"""
    return "PAR" + chr(0x2080 + n)


class MultiBandPhotosyntheticallyActiveRadiation:
    """`MultiBandPhotosyntheticallyActiveRadiation(; grid, bands, …)` — multi_band.jl:80-134."""

    def __init__(self, grid: RectilinearGrid, bands: Sequence = ((400, 500), (500, 600), (600, 700)),
                 base_bands=MOREL_λ, base_water_attenuation_coefficient=MOREL_kʷ,
                 base_chlorophyll_exponent=MOREL_e, base_chlorophyll_attenuation_coefficient=MOREL_χ,
                 field_names=None, surface_PAR=default_surface_PAR, discrete_form=False, parameters=None,
                 surface_PAR_division=None):
        self.grid = grid
        nb = len(bands)
        if not 1 <= nb <= _lib.OBM_MAX_BANDS:
            raise ValueError(f"1..{_lib.OBM_MAX_BANDS} bands supported, got {nb}")
        self.bands = tuple(bands)
        base_bands = np.asarray(base_bands, dtype=np.float64)
        kw, e, chi = [], [], []
        for lo, hi in bands:
            idx1 = int(np.nonzero(base_bands <= lo)[0][-1])  # findlast
            idx2 = int(np.nonzero(base_bands <= hi)[0][-1])
            kw.append(numerical_mean(base_bands, base_water_attenuation_coefficient, idx1, idx2))
            e.append(numerical_mean(base_bands, base_chlorophyll_exponent, idx1, idx2))
            chi.append(numerical_mean(base_bands, base_chlorophyll_attenuation_coefficient, idx1, idx2))
        self.water_attenuation_coefficient = kw
        self.chlorophyll_exponent = e
        self.chlorophyll_attenuation_coefficient = chi
        if surface_PAR_division is None:
            surface_PAR_division = [1 / nb] * nb
        if sum(surface_PAR_division) != 1:
            raise ValueError("surface_PAR_division does not sum to 1")  # multi_band.jl:106
        self.surface_PAR_division = list(surface_PAR_division)
        self.field_names = tuple(field_names) if field_names else tuple(par_symbol(n) for n in range(1, nb + 1))
        self.fields = {n: CenterField(grid, n) for n in self.field_names}
        self.total = CenterField(grid, "PAR")  # `sum(fields)` multi_band.jl:120, materialised by the kernel
        self.surface_PAR, self.discrete_form, self.parameters = surface_PAR, discrete_form, parameters

    def c_params(self) -> _lib.obm_multiband_params:
        p = _lib.obm_multiband_params()
        p.nbands = len(self.bands)
        for n in range(p.nbands):
            p.water_attenuation_coefficient[n] = self.water_attenuation_coefficient[n]
            p.chlorophyll_exponent[n] = self.chlorophyll_exponent[n]
            p.chlorophyll_attenuation_coefficient[n] = self.chlorophyll_attenuation_coefficient[n]
            p.surface_PAR_division[n] = self.surface_PAR_division[n]
        return p

    supports_column_state = True  # can produce PISCES' zₑᵤ and mixed-layer mean PAR in the same launch

    def update_biogeochemical_state(self, model, stream: Optional[int] = None, column_state=None):
        """multi_band.jl:165-185 — all bands in one launch; Chl = `chlorophyll(bgc, model)`.
        `column_state = (mixed_layer_depth, cutoff, euphotic_depth, mean_mixed_layer_light)`: also leave PISCES' two
        column diagnostics of the total PAR (obm_par_multiband_column_state)."""
        chl_a, chl_b, scale = model.biogeochemistry.chlorophyll(model)
        require_cuda(chl_a, chl_b, self.total)
        sfield, sconst = evaluate_surface_PAR(self.surface_PAR, self.grid, model.clock, self.discrete_form,
                                              self.parameters, model.tracers)
        cg, p = self.grid.c_grid(), self.c_params()
        bands = _lib.pointer_table([self.fields[n].ptr for n in self.field_names])
        s = stream if stream is not None else current_stream_ptr(self.grid.device)
        if column_state is not None:
            zmxl, cutoff, zeu, mean = column_state
            require_cuda(zmxl, zeu, mean)
            rc = _lib.load().obm_par_multiband_column_state(
                C.byref(cg), C.byref(p), chl_a.ptr, chl_b.ptr if chl_b else None, float(scale), sfield.ptr if sfield else None,
                sconst, bands, self.total.ptr, zmxl.ptr, float(cutoff), zeu.ptr, mean.ptr, s)
            _lib.check(rc, "obm_par_multiband_column_state")
            return
        rc = _lib.load().obm_par_multiband(C.byref(cg), C.byref(p), chl_a.ptr, chl_b.ptr if chl_b else None,
                                           float(scale), sfield.ptr if sfield else None, sconst, bands,
                                           self.total.ptr, s)
        _lib.check(rc, "obm_par_multiband")

    def biogeochemical_auxiliary_fields(self):
        return {"PAR": self.total, **self.fields}  # multi_band.jl:192-193

    def summary(self):
        return f"Multi band light attenuation model with {len(self.fields)} bands {self.field_names}"


class PrescribedPhotosyntheticallyActiveRadiation:
    """`PrescribedPhotosyntheticallyActiveRadiation(fields)` — src/Light/prescribed.jl: the PAR
    field(s) are supplied (and updated) by the user; nothing to compute."""

    def __init__(self, fields, field_names=None):
        if isinstance(fields, Field):
            fields = {"PAR": fields}
        elif not isinstance(fields, dict):
            fields = dict(zip(field_names or [par_symbol(n + 1) for n in range(len(fields))], fields))
        self.fields = fields

    def update_biogeochemical_state(self, model, stream=None):
        return None

    def biogeochemical_auxiliary_fields(self):
        return dict(self.fields)

    def summary(self):
        return "Prescribed PAR"


def compute_euphotic_depth(euphotic_depth: Field, PAR: Field, cutoff: float = 1 / 1000, stream=None):
    """`compute_euphotic_depth!(euphotic_depth, PAR, cutoff)` — compute_euphotic_depth.jl:31-40."""
    require_cuda(euphotic_depth, PAR)
    grid = PAR.grid
    cg = grid.c_grid()
    s = stream if stream is not None else current_stream_ptr(grid.device)
    rc = _lib.load().obm_euphotic_depth(C.byref(cg), PAR.ptr, float(cutoff), euphotic_depth.ptr, s)
    _lib.check(rc, "obm_euphotic_depth")


def compute_mixed_layer_mean(mean: Field, mixed_layer_depth: Field, C_field, grid: RectilinearGrid, stream=None):
    """`compute_mixed_layer_mean!(Cₘₓₗ, mixed_layer_depth, C, grid)` —
    PISCES/mean_mixed_layer_properties.jl:10-49.  `C_field` may be a Field or a number (ConstantField)."""
    require_cuda(mean, mixed_layer_depth)
    cg = grid.c_grid()
    s = stream if stream is not None else current_stream_ptr(grid.device)
    is_field = isinstance(C_field, Field)
    rc = _lib.load().obm_mixed_layer_mean(C.byref(cg), mixed_layer_depth.ptr, C_field.ptr if is_field else None,
                                          0.0 if is_field else float(C_field), mean.ptr, s)
    _lib.check(rc, "obm_mixed_layer_mean")
