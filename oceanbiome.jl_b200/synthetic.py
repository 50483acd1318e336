"""Deterministic synthetic tracer fields: the value of field `f` at interior linear
index `n` is lo + (hi − lo)·u (or log-uniform), u = (splitmix64(seed ⊕ f·0x9E3779B97F4A7C15 ⊕ n) >> 11)·2⁻⁵³,
seed = 20260117.  Identical bit-for-bit in numpy (host, for the oracle) and torch (device, for the
kernels), so any sub-sample of a large GPU field can be regenerated on the host for parity checks.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

SEED = 20260117
_GOLD = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_MASK = (1 << 64) - 1


def field_id(name: str) -> int:
    return zlib.crc32(name.encode("utf-8")) + 1


def _start(fid: int, seed: int) -> int:
    return (seed ^ ((fid * _GOLD) & _MASK)) & _MASK


def uniform_numpy(fid: int, start: int, count: int, seed: int = SEED) -> np.ndarray:
    """u for linear indices start … start+count-1."""
    with np.errstate(over="ignore"):
        n = np.arange(start, start + count, dtype=np.uint64)
        z = np.uint64(_start(fid, seed)) ^ n
        z = z + np.uint64(_GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _s64(x: int) -> int:
    x &= _MASK
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    """logical right shift on int64"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def uniform_torch(fid: int, start: int, count: int, device, seed: int = SEED) -> torch.Tensor:
    n = torch.arange(start, start + count, dtype=torch.int64, device=device)
    z = n ^ _s64(_start(fid, seed))
    z = z + _s64(_GOLD)
    z = (z ^ _lsr(z, 30)) * _s64(_M1)
    z = (z ^ _lsr(z, 27)) * _s64(_M2)
    z = z ^ _lsr(z, 31)
    return _lsr(z, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def _shape(u, lo, hi, log):
    if log:
        if isinstance(u, np.ndarray):
            return np.exp(np.log(lo) + (np.log(hi) - np.log(lo)) * u)
        # device exp may differ from libm by an ulp: parity tests that need bit-identical host copies
        # either fill on the host (fill_numpy) and upload, or download the device field
        return torch.exp(float(np.log(lo)) + float(np.log(hi) - np.log(lo)) * u)
    return lo + (hi - lo) * u


def fill_numpy(parent: np.ndarray, grid, name: str, lo: float, hi: float, log: bool = False, seed: int = SEED,
               rows=None):
    """Fill the interior of a halo'd parent array (linear index n = i + Nx (j + Ny k)), zero-gradient halos.
    `rows = (j0, Ny_global)`: `grid` is a y-slab of a global grid of Ny_global rows starting at global row j0 — the
    values are those rows of the GLOBAL field (n = i + Nx ((j0 + j) + Ny_global k))."""
    n3 = 1 if parent.shape[0] == 1 else grid.Nz
    fid = field_id(name)
    if rows is None:
        u = uniform_numpy(fid, 0, grid.Nx * grid.Ny * n3, seed)
        grid.interior(parent)[...] = _shape(u, lo, hi, log).reshape(n3, grid.Ny, grid.Nx)
    else:
        j0, nyg = rows
        for k in range(n3):
            u = uniform_numpy(fid, grid.Nx * (j0 + nyg * k), grid.Nx * grid.Ny, seed)
            grid.interior(parent)[k] = _shape(u, lo, hi, log).reshape(grid.Ny, grid.Nx)
    _halos(parent, grid)
    return parent


def fill_torch(field, name: str, lo: float, hi: float, log: bool = False, seed: int = SEED, chunk: int = 1 << 26,
               rows=None):
    """Same values as fill_numpy, generated on the device chunk by chunk (no host round trip unless log).
    `rows = (j0, Ny_global)` as in `fill_numpy`: this rank's rows of the global field (multi-GPU slabs)."""
    grid = field.grid
    interior = field.interior
    n3 = interior.shape[0]
    plane = grid.Nx * grid.Ny
    fid = field_id(name)
    if rows is not None:
        j0, nyg = rows
        for k in range(n3):
            u = uniform_torch(fid, grid.Nx * (j0 + nyg * k), plane, field.data.device, seed)
            interior[k] = _shape(u, lo, hi, log).reshape(grid.Ny, grid.Nx)
        field.fill_halos_zero_gradient()
        return field
    kper = max(1, chunk // plane)
    for k0 in range(0, n3, kper):
        k1 = min(n3, k0 + kper)
        u = uniform_torch(fid, k0 * plane, (k1 - k0) * plane, field.data.device, seed)
        interior[k0:k1] = _shape(u, lo, hi, log).reshape(k1 - k0, grid.Ny, grid.Nx)
    field.fill_halos_zero_gradient()
    return field


def _halos(d, g):
    if g.Hx:
        d[..., :g.Hx] = d[..., g.Hx:g.Hx + 1]
        d[..., g.Hx + g.Nx:] = d[..., g.Hx + g.Nx - 1:g.Hx + g.Nx]
    if g.Hy:
        d[:, :g.Hy] = d[:, g.Hy:g.Hy + 1]
        d[:, g.Hy + g.Ny:] = d[:, g.Hy + g.Ny - 1:g.Hy + g.Ny]
    if g.Hz and d.shape[0] != 1:
        d[:g.Hz] = d[g.Hz:g.Hz + 1]
        d[g.Hz + g.Nz:] = d[g.Hz + g.Nz - 1:g.Hz + g.Nz]


# field ranges of the BASELINE.json configs: name → (lo, hi, log)
RANGES_NPZD = {"N": (0.5, 4.5, False), "P": (0.01, 0.03, False), "Z": (0.01, 0.03, False), "D": (0.0, 0.1, False),
               "T": (8.9, 9.1, False)}
RANGES_LOBSTER = {"NO₃": (0.0, 12.0, False), "NH₄": (1e-3, 1.0, True), "P": (0.005, 0.5, True), "Z": (0.005, 0.5, True),
                  "sPOM": (0.0, 1.0, False), "bPOM": (0.0, 1.0, False), "DOM": (0.0, 1.0, False),
                  "sPON": (0.0, 1.0, False), "bPON": (0.0, 1.0, False), "DON": (0.0, 1.0, False),
                  "sPOC": (0.0, 6.56, False), "bPOC": (0.0, 6.56, False), "DOC": (0.0, 6.56, False),
                  "Fe": (1e-5, 1e-3, True), "N": (0.5, 12.0, False), "D": (0.0, 1.0, False), "T": (2.0, 28.0, False),
                  "DIC": (2000.0, 2300.0, False), "Alk": (2300.0, 2500.0, False), "O₂": (150.0, 350.0, False)}
RANGES_CARBON = {"T": (-2.0, 35.0, False), "S": (20.0, 40.0, False), "DIC": (1800.0, 2400.0, False),
                 "P": (0.0, 400.0, False), "Si": (0.0, 150.0, False), "PO₄": (0.0, 3.0, False)}


def lobster_range(name: str):
    base = name.rstrip("0123456789") if name[:3] in ("DIC", "Alk") else name
    return RANGES_LOBSTER[base]
