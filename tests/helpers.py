"""Shared test helpers: synthetic model states on device + host, and the parity metric."""
import numpy as np
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import synthetic

import pyoracle

# Stated tolerances (BASELINE.json north_star / SURVEY §8c)
RTOL_TENDENCY = 1e-12   # scale-aware relative tolerance on tendencies and PAR
ATOL_PH = 1e-10         # absolute on pH
RTOL_CARBON = 1e-10     # relative on pCO2 / fCO2 / Ω / CO3 (they inherit H)


def to_host(field) -> np.ndarray:
    return np.ascontiguousarray(field.data.detach().cpu().numpy())


def scale_aware_error(a: np.ndarray, b: np.ndarray, scale: np.ndarray) -> float:
    """max |a − b| / max(|b|, S) with S the magnitude of the additive terms forming the value
    (SURVEY §8c 'parity metric').  NaN positions must match exactly."""
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN patterns differ"
    m = ~np.isnan(b)
    den = np.maximum(np.abs(b[m]), scale[m] if isinstance(scale, np.ndarray) else scale)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(np.abs(a[m] - b[m]) / den)) if m.any() else 0.0


def synthetic_state(grid, names, ranges, device_fields=True):
    """dict name → Field filled with the deterministic synthetic values (device), plus host copies."""
    dev, host = {}, {}
    og = pyoracle.Grid.like(grid)
    for n in names:
        lo, hi, log = ranges(n) if callable(ranges) else ranges[n]
        host[n] = synthetic.fill_numpy(np.zeros(og.parent_shape), og, n, lo, hi, log)
        if device_fields:
            dev[n] = ob.CenterField(grid, n)
            dev[n].data.copy_(torch.from_numpy(host[n]))
    return dev, host, og
