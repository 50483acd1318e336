"""`ScaleNegativeTracers` / `ZeroNegativeTracers` — host-side mirror of
src/Utils/negative_tracers.jl:22-130.  A tuple of scalers (one per conserved group, as the reference
builds for PISCES / LOBSTER+carbonate) is fused into ONE launch of obm_scale_negative_tracers by
`update_biogeochemical_state(model, modifiers)` (biogeochemistry.py)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

from . import _lib
from .grids import current_stream_ptr, require_cuda


class ScaleNegativeTracers:
    """`ScaleNegativeTracers(tracers; scalefactors = ones(length(tracers)), invalid_fill_value = NaN, warn = false)`"""

    def __init__(self, tracers: Sequence[str], scalefactors: Optional[Sequence[float]] = None,
                 invalid_fill_value: float = float("nan"), warn: bool = False):
        if warn:
            raise NotImplementedError("Warning not currently implemented")  # negative_tracers.jl:45
        tracers = tuple(tracers)
        scalefactors = tuple(float(s) for s in (scalefactors if scalefactors is not None else [1.0] * len(tracers)))
        if len(scalefactors) != len(tracers):
            raise ValueError("Incorrect number of scale factors provided")  # :83-85
        if len(tracers) > _lib.OBM_MAX_GROUP_SIZE:
            raise ValueError(f"at most {_lib.OBM_MAX_GROUP_SIZE} tracers per group")
        self.tracers, self.scalefactors, self.invalid_fill_value = tracers, scalefactors, float(invalid_fill_value)

    @staticmethod
    def from_biogeochemistry(bgc, grid=None, invalid_fill_value=float("nan"), warn=False):
        """`ScaleNegativeTracers(bgc, grid; …)` :97-130 — one scaler per conserved group."""
        groups = bgc.conserved_tracers()
        if groups and isinstance(groups[0], str):
            return ScaleNegativeTracers(groups, invalid_fill_value=invalid_fill_value, warn=warn)
        out = []
        for g in groups:
            if isinstance(g, dict):
                out.append(ScaleNegativeTracers(g["tracers"], g["scalefactors"], invalid_fill_value, warn))
            else:
                out.append(ScaleNegativeTracers(g, None, invalid_fill_value, warn))
        return tuple(out)

    def update_biogeochemical_state(self, model, stream=None):
        apply_scalers(model, (self,), stream)

    def summary(self):
        return f"Mass conserving negative scaling of {self.tracers}"


def apply_scalers(model, scalers: Sequence[ScaleNegativeTracers], stream: Optional[int] = None, calcite=None) -> bool:
    """All groups in one launch, applied in order (OceanBioME.jl:169 applies modifiers in tuple order).
    `calcite`: optional `(carbon_chemistry, T, S, DIC, Alk, Si, Ω, [H⁺] state or None)` — the PISCES calcite-saturation
    solve of the same cells rides in the same launch (obm_scale_negative_tracers_calcite_saturation).  Returns
    whether it did."""
    names = []
    for sc in scalers:
        for t in sc.tracers:
            if t not in names:
                names.append(t)
    if len(names) > _lib.OBM_MAX_SCALE_TRACERS or len(scalers) > _lib.OBM_MAX_SCALE_GROUPS:
        raise ValueError("too many tracers / groups for one fused launch")
    fills = {sc.invalid_fill_value if sc.invalid_fill_value == sc.invalid_fill_value else "nan" for sc in scalers}
    if len(fills) > 1:  # different fill values cannot share a launch: fall back to one launch per group
        for sc in scalers:
            apply_scalers(model, (sc,), stream)
        return False
    fields = [model.tracers[n] for n in names]
    require_cuda(*fields)
    groups = (_lib.obm_scale_group * len(scalers))()
    for q, sc in enumerate(scalers):
        groups[q].n = len(sc.tracers)
        for m, (t, f) in enumerate(zip(sc.tracers, sc.scalefactors)):
            groups[q].index[m] = names.index(t)
            groups[q].scalefactor[m] = f
    grid = model.grid
    cg = grid.c_grid()
    s = stream if stream is not None else current_stream_ptr(grid.device)
    if calcite is not None:
        cc, T, S, DIC, Alk, Si, Omega, state = calcite
        require_cuda(T, S, DIC, Alk, Si, Omega)
        p = cc.c_params()
        rc = _lib.load().obm_scale_negative_tracers_calcite_saturation(
            C.byref(cg), len(names), _lib.pointer_table([f.ptr for f in fields]), len(scalers), groups,
            scalers[0].invalid_fill_value, C.byref(p), T.ptr, S.ptr, DIC.ptr, Alk.ptr, Si.ptr, Omega.ptr,
            state.ptr if state is not None else None, s)
        _lib.check(rc, "obm_scale_negative_tracers_calcite_saturation")
        return True
    rc = _lib.load().obm_scale_negative_tracers(C.byref(cg), len(names), _lib.pointer_table([f.ptr for f in fields]),
                                                len(scalers), groups, scalers[0].invalid_fill_value, s)
    _lib.check(rc, "obm_scale_negative_tracers")
    return False


class ZeroNegativeTracers:
    """`ZeroNegativeTracers(; exclude = ())` — negative_tracers.jl:22-32 (does not conserve mass)."""

    def __init__(self, exclude: Sequence[str] = ()):
        self.exclude = tuple(exclude)

    def update_biogeochemical_state(self, model, stream=None):
        fields = [f for n, f in model.tracers.items() if n not in self.exclude]
        if not fields:
            return
        require_cuda(*fields)
        s = stream if stream is not None else current_stream_ptr(model.grid.device)
        lib = _lib.load()
        for c in range(0, len(fields), _lib.OBM_MAX_SCALE_TRACERS):
            chunk = fields[c:c + _lib.OBM_MAX_SCALE_TRACERS]
            rc = lib.obm_zero_negative_tracers(chunk[0].data.numel(), len(chunk),
                                               _lib.pointer_table([f.ptr for f in chunk]), s)
            _lib.check(rc, "obm_zero_negative_tracers")

    def summary(self):
        return f"Zero negative tracers (excluding {self.exclude})"
