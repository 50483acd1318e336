"""GPU parity: fused ScaleNegativeTracers / ZeroNegativeTracers / inventory against the oracle.
Decisions are exact (the sums t, p are formed with the reference's operation sequence, no FMA contraction), so
NaN and zero patterns are bit-identical; rescaled values agree to ≤ 2 ulp per applied group (one division per group: v·(t/p) vs
v·t/p; untouched groups are left alone where the reference rewrites v·t/t).  Plus the reference's own expected
outcomes (test_utils.jl, test_PISCES.jl:129-168).  ZeroNegativeTracers and the inventory order are bit-exact."""
import math

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import synthetic
from helpers import synthetic_state
from test_oracle_negs import PISCES_GROUPS

pytestmark = pytest.mark.gpu

PISCES_TRACERS = ("P", "D", "Z", "M", "PChl", "DChl", "PFe", "DFe", "DSi", "DOC", "POC", "GOC", "SFe", "BFe", "PSi", "NO₃",
                  "NH₄", "PO₄", "Fe", "Si", "CaCO₃", "DIC", "Alk", "O₂")


class M:
    def __init__(self, grid, tracers):
        self.grid, self.tracers, self.clock = grid, tracers, ob.Clock()


def pisces_scalers():
    groups = list(PISCES_GROUPS)
    groups[1] = (groups[1][0], (1, 1, 0.01, 0.015, 1, 1, 1))
    return groups, tuple(ob.ScaleNegativeTracers(t, s) for t, s in groups)


def test_pisces_groups_match_oracle(cuda, oracle):
    grid = ob.RectilinearGrid(size=(41, 6, 11), extent=(41, 6, 110), device=cuda)
    ranges = {n: (-0.3, 1.0, False) for n in PISCES_TRACERS}  # ~23 % negative entries
    dev, host, og = synthetic_state(grid, PISCES_TRACERS, ranges)
    for n in ("P", "Si"):  # sprinkle non-finite values: they are rescaled like positives (negative_tracers.jl:271)
        host[n][og.Hz + 2, og.Hy + 1, og.Hx + 5] = np.inf
        host[n][og.Hz + 3, og.Hy + 2, og.Hx + 7] = np.nan
        dev[n].data.copy_(torch.from_numpy(host[n]))
    groups, scalers = pisces_scalers()
    before = {n: host[n].copy() for n in PISCES_TRACERS}
    bgc = ob.Biogeochemistry(None, modifiers=scalers)
    ob.biogeochemistry._update_modifiers(M(grid, dev), scalers, None)
    oracle.scale_negative_tracers(og, [host[n] for n in PISCES_TRACERS], oracle.make_groups(PISCES_TRACERS, groups))
    changed = 0
    for n in PISCES_TRACERS:
        got = dev[n].data.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(host[n])), n          # NaN pattern: exact
        assert np.array_equal(got == 0, host[n] == 0), n                    # zero pattern: exact
        assert np.array_equal(np.isinf(got), np.isinf(host[n])), n
        fin = np.isfinite(host[n])
        # ≤ 2 ulp per applied group; a later group whose total is a near-cancellation (t ≈ 0 from O(1) terms) amplifies
        # those ulps, so the bound is scale-aware like the tendency metric: 1e-12 of max(|value|, magnitude of the inputs = 1)
        assert np.all(np.abs(got[fin] - host[n][fin]) <= 1e-12 * np.maximum(np.abs(host[n][fin]), 1.0)), n
        full = got.copy()
        og.interior(full)[...] = og.interior(before[n])
        assert np.array_equal(full, before[n], equal_nan=True), n           # halos untouched
        changed += int(not np.array_equal(host[n], before[n], equal_nan=True))
    assert changed >= 20
    assert np.isnan(og.interior(host["DSi"])).any() and (og.interior(host["P"]) == 0).any()


def test_reference_expected_outcomes(cuda):
    grid = ob.RectilinearGrid(size=(1, 1, 1), extent=(1, 1, 1), device=cuda)
    # test_utils.jl:7-22
    model = ob.BiogeochemicalModel(grid, ob.NPZD(grid, scale_negatives=True, surface_photosynthetically_active_radiation=100.0))
    model.set(N=2, P=-1)
    model.time_step(1e-10)
    assert math.isclose(model.tracers["N"].interior.item(), 1, rel_tol=1e-8) and abs(model.tracers["P"].interior.item()) < 1e-12
    # test_utils.jl:24-40
    model = ob.BiogeochemicalModel(grid, ob.NPZD(grid, modifiers=ob.ZeroNegativeTracers(exclude=("Z",)), surface_photosynthetically_active_radiation=100.0))
    model.set(N=2, P=-1, Z=-1)
    model.time_step(1e-10)
    v = {n: model.tracers[n].interior.item() for n in "NPZ"}
    assert math.isclose(v["N"], 2, rel_tol=1e-8) and abs(v["P"]) < 1e-12 and math.isclose(v["Z"], -1, rel_tol=1e-8)
    # test_PISCES.jl:147-167
    tr = {n: ob.CenterField(grid, n) for n in PISCES_TRACERS}
    for n in ("D", "Z", "M", "DOC", "POC", "GOC", "DIC", "CaCO₃", "PO₄"):
        tr[n].set(1.0)
    tr["P"].set(-1.0)
    _, scalers = pisces_scalers()
    m = M(grid, tr)
    ob.biogeochemistry._update_modifiers(m, scalers, None)
    assert tr["P"].interior.item() == 0
    assert all(math.isclose(tr[n].interior.item(), 7 / 8, rel_tol=1e-8) for n in ("D", "Z", "M", "DOC", "POC", "GOC", "DIC", "CaCO₃"))
    assert tr["PO₄"].interior.item() == 1
    tr["Si"].set(-1.0)
    tr["DSi"].set(0.1)
    ob.biogeochemistry._update_modifiers(m, scalers, None)
    assert math.isnan(tr["DSi"].interior.item())
    tr["Fe"].set(-1.0)
    tr["Z"].set(1000.0)
    tr["M"].set(0.0)
    ob.biogeochemistry._update_modifiers(m, scalers, None)
    assert tr["Fe"].interior.item() == 0 and math.isclose(tr["Z"].interior.item(), 900, rel_tol=1e-8)


def test_absorbed_negative_member_is_still_zeroed(cuda, oracle):
    """A negative member far below the ulp of the group's positive sum (the usual overshoot: DIC ≈ 2e3, plankton ≈ −1e-15)
    leaves t == p bit for bit; the reference still writes 0 for every non-positive member (negative_tracers.jl:268-274)
    and leaves the positives as they are (t / p = 1).  Also −0.0 → +0.0, as `ifelse(…, 0)` gives."""
    grid = ob.RectilinearGrid(size=(8, 2, 3), extent=(8, 2, 3), device=cuda)
    names = ("P", "Z", "DIC")
    dev, host, og = synthetic_state(grid, names, {"P": (0.01, 0.5, False), "Z": (0.01, 0.5, False), "DIC": (2000.0, 2300.0, False)})
    for (k, j, i), v in {(0, 0, 1): -1e-14, (1, 1, 3): -5e-324, (2, 0, 5): -0.0, (2, 1, 7): -1e-13}.items():
        host["P"][og.Hz + k, og.Hy + j, og.Hx + i] = v
    dev["P"].data.copy_(torch.from_numpy(host["P"]))
    groups = [(names, (6.56, 6.56, 1.0))]
    before = {n: host[n].copy() for n in names}
    scalers = tuple(ob.ScaleNegativeTracers(t, s) for t, s in groups)
    ob.biogeochemistry._update_modifiers(M(grid, dev), scalers, None)
    oracle.scale_negative_tracers(og, [host[n] for n in names], oracle.make_groups(names, groups))
    gotP = og.interior(dev["P"].data.cpu().numpy())
    assert bool((gotP >= 0).all()) and int((gotP == 0).sum()) == 4 and not np.signbit(gotP).any()
    for n in names:
        got = dev[n].data.cpu().numpy()
        assert np.array_equal(got == 0, host[n] == 0), n
        assert np.all(np.abs(got - host[n]) <= 5e-16 * np.abs(host[n])), n  # (v·t)/p vs v·(t/p): ≤ 2 ulp
        touched = og.interior(before["P"]) <= 0
        assert np.array_equal(og.interior(got)[~touched], og.interior(before[n])[~touched]), n  # clean cells: bit for bit


def test_immersed_cells_are_left_alone(cuda, oracle):
    """Grid-fitted bottom: `grid.immersed(bottom_height)` = `ImmersedBoundaryGrid(grid, GridFittedBottom(h))`; the scaling
    kernels skip the cells below the bottom-most active one like the reference's `!immersed_cell` guard
    (negative_tracers.jl:194,253) — against the oracle, bit for bit where nothing is rescaled."""
    base = ob.RectilinearGrid(size=(29, 7, 13), extent=(29, 7, 130), device=cuda)
    h = ob.Field2D(base, "bottom_height")
    synthetic.fill_torch(h, "bottom_height", -140.0, -5.0)
    grid = base.immersed(h)
    names = ("P", "Z", "NO₃", "NH₄")
    dev, host, og = synthetic_state(grid, names, {n: (-0.4, 1.0, False) for n in names})
    assert og.bottom_indices is not None and og.interior(og.bottom_indices).max() > 3
    before = {n: host[n].copy() for n in names}
    groups = [(names, (1, 1, 1, 1))]
    scalers = tuple(ob.ScaleNegativeTracers(t, s) for t, s in groups)
    ob.biogeochemistry._update_modifiers(M(grid, dev), scalers, None)
    oracle.scale_negative_tracers(og, [host[n] for n in names], oracle.make_groups(names, groups))
    kb = og.interior(og.bottom_indices)[0] - 1                       # 0-based bottom-most active k per column
    dry = np.arange(grid.Nz).reshape(-1, 1, 1) < kb[None]            # immersed cells
    assert dry.any() and (~dry).any()
    for n in names:
        got = dev[n].data.cpu().numpy()
        assert np.array_equal(og.interior(got)[dry], og.interior(before[n])[dry]), n   # untouched, negatives included
        assert not (og.interior(got)[~dry] < 0).any(), n               # (a group whose total is negative is NaN-filled)
        assert np.array_equal(got == 0, host[n] == 0) and np.array_equal(np.isnan(got), np.isnan(host[n])), n
        fin = np.isfinite(host[n])
        assert np.all(np.abs(got[fin] - host[n][fin]) <= 1e-12 * np.maximum(np.abs(host[n][fin]), 1.0)), n
    assert (og.interior(dev["P"].data.cpu().numpy())[dry] < 0).any()


def test_zero_negative_bit_exact(cuda, oracle):
    grid = ob.RectilinearGrid(size=(13, 5, 7), extent=(13, 5, 7), device=cuda)
    dev, host, og = synthetic_state(grid, ["A", "B", "C"], {n: (-1.0, 1.0, False) for n in "ABC"})
    host["B"][4, 4, 4] = np.nan
    dev["B"].data.copy_(torch.from_numpy(host["B"]))
    ob.ZeroNegativeTracers(exclude=("C",)).update_biogeochemical_state(M(grid, dev))
    oracle.zero_negative_tracers([host["A"], host["B"]])
    for n in "ABC":
        assert np.array_equal(dev[n].data.cpu().numpy(), host[n], equal_nan=True)


def test_inventory_matches_oracle(cuda, oracle):
    from oceanbiome_b200.distributed import tracer_inventory
    grid = ob.RectilinearGrid(size=(61, 10, 17), x=(0, 61), y=(0, 10), z=lambda k: -100 * (1 - ((k - 1) / 17) ** 1.5), device=cuda)
    names = ("P", "Z", "NO₃", "NH₄", "sPOM", "bPOM", "DOM", "DIC")
    dev, host, og = synthetic_state(grid, names, synthetic.lobster_range)
    groups = [(names[:7], (1,) * 7), (("P", "Z", "DIC", "sPOM", "bPOM", "DOM"), (1.1 * 6.56, 6.56, 1, 6.56, 6.56, 6.56))]
    got = tracer_inventory(grid, dev, groups).cpu().numpy()
    vol = np.zeros(og.parent_shape)
    og.interior(vol)[...] = (grid.dz * grid.dx * grid.dy).reshape(-1, 1, 1)
    want = oracle.inventory(og, [host[n] for n in names], oracle.make_groups(names, groups), cell_volume=vol)
    np.testing.assert_allclose(got, want, rtol=1e-13)
    # deterministic: run-to-run bit-identical
    again = tracer_inventory(grid, dev, groups).cpu().numpy()
    assert np.array_equal(got, again)
