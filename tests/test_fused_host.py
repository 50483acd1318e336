"""CPU: the fused PISCES kernel's OWN arithmetic (csrc/pisces_cell.cuh, csrc/carbon_chemistry.cuh — the headers the CUDA
kernels are built from) compiled for the host (bench_ref/fused_host.cpp) and compared with the oracle.  This checks the
kernel source's operation order, parameter plumbing and branch structure WITHOUT a GPU — both arithmetic policies: the
exact pass (IEEE division, NaN-propagating min / max) and the fast pass (selects; its reciprocal is IEEE 1/x on the
host, where the device uses a ≤ 1 ulp sequence) — at the stated per-tendency tolerance.  The GPU tests repeat the
comparison with the device arithmetic."""
import os
import sys

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import pisces, synthetic
from helpers import RTOL_TENDENCY, tendency_parity

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench_ref"))


@pytest.fixture(scope="module")
def fused():
    import fused_host
    fused_host.build()
    return fused_host


def state(oracle, size=(24, 5, 12), t=0.37 * 365 * 86400.0):
    grid = ob.RectilinearGrid(size=size, extent=(1e4, 1e3, 400.0), device="cpu")
    og = oracle.Grid.like(grid)
    bgc = ob.PISCES(grid, scale_negatives=True)
    u = bgc.underlying_biogeochemistry
    host = {n: synthetic.fill_numpy(np.zeros(og.parent_shape), og, n, *pisces.synthetic_range(n)) for n in pisces.TRACERS}
    zmxl = synthetic.fill_numpy(np.zeros(og.plane_shape), og, "zₘₓₗ", -150.0, -10.0)
    kappa = synthetic.fill_numpy(np.zeros(og.plane_shape), og, "κ̄", 1e-4, 1e-2, True)
    u.mixed_layer_depth.data.copy_(torch.from_numpy(zmxl))
    u.euphotic_depth.data.fill_(-60.0)
    wPOC = np.ascontiguousarray(u.sinking_velocities["POC"].data.numpy())
    wGOC = np.ascontiguousarray(pisces.DepthDependantSinkingSpeed().face_field(grid, u.mixed_layer_depth, u.euphotic_depth).data.numpy())
    return grid, og, bgc, u, host, zmxl, kappa, wPOC, wGOC, t


@pytest.mark.parametrize("exact", [True, False])
def test_kernel_cell_arithmetic_matches_oracle_on_the_host(oracle, fused, exact):
    grid, og, bgc, u, host, zmxl, kappa, wPOC, wGOC, t = state(oracle)
    bands, total = oracle.par_multiband(og, bgc.light_attenuation.c_params(), host["PChl"], host["DChl"], 1.0, 100.0)
    zeu = oracle.euphotic_depth(og, total)
    mean = oracle.mixed_layer_mean(og, zmxl, total)
    Om = synthetic.fill_numpy(np.zeros(og.parent_shape), og, "Ω", 0.3, 4.0)  # both sides of Ω = 1
    aux = {"PAR1": bands[0], "PAR2": bands[1], "PAR3": bands[2], "PAR": total, "Omega": Om, "wPOC": wPOC, "wGOC": wGOC,
           "mixed_layer_depth_xy": zmxl, "euphotic_depth_xy": zeu, "mean_mixed_layer_vertical_diffusivity_xy": kappa,
           "mean_mixed_layer_light_xy": mean}
    tr = [host[n] for n in pisces.TRACERS]
    p = u.c_params(t)
    want = oracle.pisces_tendencies(og, p, tr, aux)
    S = oracle.pisces_tendency_scales(og, p, tr, aux)
    got = fused.pisces_tendencies(og, p, tr, aux, exact=exact)
    worst = 0.0
    for n in range(24):
        err, rel_max, rel_p = tendency_parity(og.interior(got[n]), og.interior(want[n]), og.interior(S[n]))
        assert err <= RTOL_TENDENCY and rel_p <= 1e-12, (pisces.TRACERS[n], err, rel_max, rel_p)
        worst = max(worst, err)
    print(f"[parity] host build of the kernel's cell arithmetic (exact={exact}): scale-aware max {worst:.2e}")
    assert got[24] is None and got[25] is None  # T, S: zero(grid)


def test_fused_prologue_and_light_match_oracle_on_the_host(oracle, fused):
    """The other two launches of the stage as the kernels fuse them (all groups + Ω of the rescaled cell; all bands + zₑᵤ +
    PAR̄ in one column pass) against the oracle's one-pass-per-group / per-band structure."""
    grid, og, bgc, u, host, zmxl, kappa, wPOC, wGOC, t = state(oracle, size=(16, 4, 20))
    host["DOC"][og.Hz + 3, og.Hy + 1, og.Hx:og.Hx + 6] = -5.0   # carbon group rescales DIC in these cells
    host["Si"][og.Hz + 5, og.Hy + 2, og.Hx + 4] = -0.3
    groups = [(m.tracers, m.scalefactors) for m in bgc.modifiers]
    snames = []
    for tn, _ in groups:
        snames += [x for x in tn if x not in snames]
    cg = oracle.make_groups(snames, groups)
    mine = {n: host[n].copy() for n in host}
    Om = fused.scale_negative_tracers_calcite_saturation(og, [mine[n] for n in snames], cg, mine["T"], mine["S"], mine["DIC"], mine["Alk"], mine["Si"])
    # the per-level tables (TEOS-10 collapsed in ζ, pressure corrections as quadratics in T_c) against the direct form
    direct = {n: host[n].copy() for n in host}
    Om_direct = fused.scale_negative_tracers_calcite_saturation(og, [direct[n] for n in snames], cg, direct["T"], direct["S"],
                                                                direct["DIC"], direct["Alk"], direct["Si"], level_tables=False)
    a, b = og.interior(Om), og.interior(Om_direct)
    both = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and float(np.max(np.abs(a[both] - b[both]) / np.abs(b[both]))) <= 1e-13
    oracle.scale_negative_tracers(og, [host[n] for n in snames], cg)
    for n in snames:
        a, b = og.interior(mine[n]), og.interior(host[n])
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a == 0, b == 0), n
        ok = np.isfinite(b)
        assert np.all(np.abs(a[ok] - b[ok]) <= 1e-12 * np.maximum(np.abs(b[ok]), 1.0)), n
    want = og.interior(oracle.calcite_saturation(og, host["T"], host["S"], host["DIC"], host["Alk"], host["Si"]))
    got = og.interior(Om)
    ok = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.max(np.abs(got[ok] - want[ok]) / np.abs(want[ok])) <= 1e-10
    bands, total, zeu, mean = fused.par_multiband_column_state(og, bgc.light_attenuation.c_params(), host["PChl"], host["DChl"], 1.0, 100.0, zmxl)
    ob_, ot = oracle.par_multiband(og, bgc.light_attenuation.c_params(), host["PChl"], host["DChl"], 1.0, 100.0)
    rel = lambda a, b: float(np.max(np.abs(og.interior(a) - og.interior(b)) / np.abs(og.interior(b))))  # noqa: E731
    assert max(rel(bands[n], ob_[n]) for n in range(3)) <= 1e-12 and rel(total, ot) <= 1e-12
    assert rel(zeu, oracle.euphotic_depth(og, ot)) <= 1e-12 and rel(mean, oracle.mixed_layer_mean(og, zmxl, ot)) <= 1e-12


def test_fp32_presolve_then_one_fp64_step_reaches_the_reference_root(oracle, fused):
    """carbon_chemistry.cuh presolve_f32: three FP32 Newton steps from the carbonate-alkalinity quadratic, then ONE FP64 step
    whose C·Δx² error is removed with the C the pre-solve delivers.  On sea-water states that single FP64 step
    (iterations = 1) must already sit on the root of the reference's damped Newton (oracle): |ΔpH| ≤ 1e-12 here against
    the stated 1e-10; over a box far outside sea water the ordinary FP64 loop takes over and the stated bound holds."""
    from oceanbiome_b200 import _lib as abi
    rng = np.random.default_rng(11)
    n = 20000
    T, S = rng.uniform(-1.5, 32, n), rng.uniform(28, 40, n)
    DIC = rng.uniform(1800, 2400, n)
    Alk = DIC * rng.uniform(1.02, 1.25, n)
    want, _ = oracle.carbon_chemistry_sweep(T, S, DIC, Alk, output=abi.CC_PH_FREE)
    for iterations in (1, 12):
        got = fused.carbon_chemistry_sweep(T, S, DIC, Alk, iterations=iterations)
        err = float(np.max(np.abs(got - want)))
        print(f"[parity] FP32 pre-solve + {iterations} FP64 step(s), sea water: max |dpH| {err:.2e}")
        assert err <= 1e-12, (iterations, err)
    for kind in (abi.CC_OMEGA_CALCITE, abi.CC_FCO2):
        w, _ = oracle.carbon_chemistry_sweep(T, S, DIC, Alk, output=kind)
        g = fused.carbon_chemistry_sweep(T, S, DIC, Alk, output=kind)
        assert float(np.max(np.abs(g - w) / np.abs(w))) <= 1e-12, kind
    T, S = rng.uniform(-2, 40, n), rng.uniform(1, 45, n)
    DIC, Alk = rng.uniform(100, 4000, n), rng.uniform(100, 4500, n)
    want, _ = oracle.carbon_chemistry_sweep(T, S, DIC, Alk, output=abi.CC_PH_FREE)
    got = fused.carbon_chemistry_sweep(T, S, DIC, Alk)
    ok = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), ok) and float(np.max(np.abs(got - want)[ok])) <= 1e-10
