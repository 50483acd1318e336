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


def tendency_parity(got: np.ndarray, want: np.ndarray, S: np.ndarray, offset: float = 0.0):
    """The stated parity metric, per tendency (BASELINE.json north_star "relative 1e-12 on tendencies"):

        err = |got − want| / max(|want|, S),   S = Σ |additive terms| of THAT tendency in that cell

    (S from the oracle's `*_tendency_scales`; in accumulate mode the pre-existing Gⁿ value `offset` is one more term).
    Also the PURE relative error |got − want| / |want − offset| over the cells where the tendency is not a near-total
    cancellation of its terms (|want − offset| ≥ 1e-3·S): (max, 99.9th percentile).  NaN positions must match exactly.
    Returns (err_max, rel_max, rel_p999)."""
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN patterns differ"
    m = np.isfinite(want) & np.isfinite(S)
    inf = ~np.isnan(want) & ~m
    assert np.array_equal(got[inf], want[inf]), "±Inf patterns differ"
    if not m.any():
        return 0.0, 0.0, 0.0
    g, w, s = got[m], want[m], S[m] + abs(offset)
    d = np.abs(g - w)
    den = np.maximum(np.abs(w), s)
    err = float(np.max(d / np.where(den == 0, 1.0, den)))
    # accumulate mode: relative to the tendency itself, |want − offset| — offset + tendency can land near zero (a
    # cancellation between the pre-existing Gⁿ and what was added, nothing to do with the kernel's arithmetic)
    big = (np.abs(w - offset) >= 1e-3 * s) & (np.abs(w - offset) > 0)
    if big.any():
        rel = d[big] / np.abs(w[big] - offset)
        return err, float(rel.max()), float(np.quantile(rel, 0.999))
    return err, 0.0, 0.0


# what the parity tests measured, appended per call to gpurun_out/parity_metrics.jsonl (read back into DESIGN §4)
def record_parity(test: str, metrics: dict):
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_metrics.jsonl"), "a") as f:
            f.write(json.dumps({"test": test, **metrics}) + "\n")
    except OSError:
        pass


RTOL_PURE_P999 = 1e-12  # 99.9th percentile of the pure relative error where |tendency| ≥ 1e-3·S
RTOL_PURE_MAX = 1e-10   # its maximum (bounded by 1e-12·S/|tendency| ≤ 1e-9 through the scale-aware metric alone)


def assert_tendency_parity(test: str, names, got: dict, want: dict, S: dict, offset: float = 0.0):
    """Assert the per-tendency metric for every tracer of `names`; print and record the numbers."""
    rows = {}
    for n in names:
        rows[n] = tendency_parity(got[n], want[n], S[n], offset)
    worst = max(r[0] for r in rows.values())
    rmax = max(r[1] for r in rows.values())
    rp = max(r[2] for r in rows.values())
    print(f"[parity] {test}: scale-aware max {worst:.2e}; pure relative max {rmax:.2e}, p99.9 {rp:.2e}")
    record_parity(test, {"scale_aware_max": worst, "pure_rel_max": rmax, "pure_rel_p999": rp,
                         "per_tracer": {n: [float(f"{x:.3e}") for x in r] for n, r in rows.items()}})
    bad = {n: r for n, r in rows.items() if r[0] > RTOL_TENDENCY or r[2] > RTOL_PURE_P999 or r[1] > RTOL_PURE_MAX}
    assert not bad, f"{test}: tendency parity (scale-aware, pure max, pure p99.9) {bad}"
    return worst, rmax, rp


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
