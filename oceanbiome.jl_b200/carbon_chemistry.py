"""`CarbonChemistry` — host-side mirror of src/Models/CarbonChemistry/carbon_chemistry.jl:66-165.

`cc(DIC=…, T=…, S=…, Alk=…, pH=None, P=None, output="fCO₂", silicate=None, phosphate=None)` works on
flat CUDA tensors of any common shape; the solve runs in csrc/carbon_chemistry.cu.  Only the default
constants of `CarbonChemistry()` are supported (they are compiled into the kernel).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .grids import Field, RectilinearGrid, current_stream_ptr

OUTPUTS = {
    "fCO₂": _lib.CC_FCO2, "fCO2": _lib.CC_FCO2,
    "pCO₂": _lib.CC_PCO2, "pCO2": _lib.CC_PCO2,
    "pHᶠ": _lib.CC_PH_FREE, "pHf": _lib.CC_PH_FREE,
    "pHᵗ": _lib.CC_PH_TOTAL, "pHt": _lib.CC_PH_TOTAL,
    "pHˢ": _lib.CC_PH_SEAWATER, "pHs": _lib.CC_PH_SEAWATER,
    "CO₃²⁻": _lib.CC_CO3, "CO3": _lib.CC_CO3,
    "Ω": _lib.CC_OMEGA_CALCITE, "omega_calcite": _lib.CC_OMEGA_CALCITE,
}


class CarbonChemistry:
    """`CarbonChemistry(FT = Float64; …)` with the defaults of carbon_chemistry.jl:66-87.

    `newton_iterations`: fixed iteration count of the branch-free ln[H⁺] Newton that replaces the
    reference's DampedNewtonRaphsonSolver (8 suffices for model states, 12 is the robust default)."""

    def __init__(self, newton_iterations: int = 12, initial_pH_guess: float = 8.0):
        self.newton_iterations = int(newton_iterations)
        self.initial_pH_guess = float(initial_pH_guess)

    def c_params(self) -> _lib.obm_carbchem_params:
        return _lib.obm_carbchem_params(self.newton_iterations, 0, self.initial_pH_guess)

    def __call__(self, *, DIC, T, S, Alk=None, pH=None, P=None, output="fCO₂", silicate=None, phosphate=None,
                 out: Optional[torch.Tensor] = None, stream: Optional[int] = None) -> torch.Tensor:
        kind = OUTPUTS[output] if isinstance(output, str) else int(output)
        DIC = _as_cuda(DIC)
        dev, n = DIC.device, DIC.numel()
        if Alk is None and pH is None:
            Alk = torch.zeros_like(DIC)  # `Alk = zero(DIC)` default, carbon_chemistry.jl:111

        def prep(x):
            if x is None:
                return None
            x = _as_cuda(x, dev)
            return x.expand_as(DIC).contiguous() if x.numel() != n else x.contiguous()

        T, S, Alk, pH, P, silicate, phosphate = (prep(x) for x in (T, S, Alk, pH, P, silicate, phosphate))
        DIC = DIC.contiguous()
        if out is None:
            out = torch.empty_like(DIC)
        p = self.c_params()
        s = stream if stream is not None else current_stream_ptr(dev)
        ptr = lambda x: x.data_ptr() if x is not None else None  # noqa: E731
        rc = _lib.load().obm_carbon_chemistry(n, C.byref(p), ptr(T), ptr(S), ptr(DIC), ptr(Alk), ptr(P),
                                              ptr(silicate), ptr(phosphate), ptr(pH), kind, out.data_ptr(), s)
        _lib.check(rc, "obm_carbon_chemistry")
        return out

    def calcite_saturation(self, grid: RectilinearGrid, T: Field, S: Field, DIC: Field, Alk: Field, Si: Field,
                           Omega: Field, stream: Optional[int] = None, state: Optional[Field] = None):
        """`compute_calcite_saturation!` — PISCES/compute_calcite_saturation.jl:9-37.
        `state`: optional field carrying [H⁺] from call to call (Newton warm start)."""
        cg, p = grid.c_grid(), self.c_params()
        s = stream if stream is not None else current_stream_ptr(grid.device)
        rc = _lib.load().obm_calcite_saturation(C.byref(cg), C.byref(p), T.ptr, S.ptr, DIC.ptr, Alk.ptr, Si.ptr,
                                                Omega.ptr, state.ptr if state is not None else None, s)
        _lib.check(rc, "obm_calcite_saturation")

    def summary(self):
        return "`CarbonChemistry` model"

    def __repr__(self):
        return "`CarbonChemistry` model which solves for pCO₂ and pH"


def _as_cuda(x, device=None):
    if not torch.is_tensor(x):
        if device is None:
            raise RuntimeError("oceanbiome.jl_b200 kernels need CUDA tensors: there is no CPU fallback")
        x = torch.as_tensor(x, dtype=torch.float64, device=device)
    if not x.is_cuda:
        raise RuntimeError("oceanbiome.jl_b200 kernels need CUDA tensors: there is no CPU fallback")
    return x.to(torch.float64)
