"""GPU: the sinking-advection kernel (obm_sinking_tendencies) against the oracle on identical inputs, and the
reference's sediment conservation property (test/test_sediments.jl:37-80) with the whole step on the device:
BGC tendencies + sinking into the sediment + RK3/Euler update, total nitrogen constant."""
import ctypes as C

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu


def host(f):
    return np.ascontiguousarray(f.data.cpu().numpy())


@pytest.mark.parametrize("scheme", [_lib.ADV_UPWIND1, _lib.ADV_CENTERED2, _lib.ADV_UPWIND3, _lib.ADV_WENO5])
@pytest.mark.parametrize("accumulate", [False, True])
def test_kernel_matches_oracle(cuda, oracle, scheme, accumulate):
    grid = ob.RectilinearGrid(size=(37, 5, 23), x=(0, 37), y=(0, 5), z=np.cumsum(np.r_[-60.0, 1.0 + np.arange(23) * 0.15]),
                              device=cuda)
    og = oracle.Grid.like(grid)
    names = ["a", "b", "c"]
    c = {n: ob.CenterField(grid, n) for n in names}
    w = {n: ob.ZFaceField(grid, "w" + n) for n in names}
    G = {n: ob.CenterField(grid, "G" + n, 0.5) for n in names}
    for q, n in enumerate(names):
        synthetic.fill_torch(c[n], n, 0.01, 3.0, log=True)
        c[n].data[:grid.Hz] = 0.7  # filled bottom halo (read by upwind w > 0 / centred at the bottom face)
        synthetic.fill_torch(w[n], "w" + n, -5e-3, 5e-3 if q == 2 else -1e-4)  # tracer c: both signs
        w[n].face_interior[grid.Nz] = 0.0  # closed top
    cg = grid.c_grid()
    tab = lambda d: _lib.pointer_table([d[n].ptr for n in names])  # noqa: E731
    hc, hw, hG = [host(c[n]) for n in names], [host(w[n]) for n in names], [host(G[n]) for n in names]
    rc = _lib.load().obm_sinking_tendencies(C.byref(cg), 3, tab(c), tab(w), tab(G), scheme, int(accumulate), None)
    _lib.check(rc, "obm_sinking_tendencies")
    oracle.sinking_tendencies(og, hc, hw, hG, scheme, accumulate)
    for n, want in zip(names, hG):
        got = host(G[n])
        scale = np.abs(og.interior(want)).max()
        # same operations; FMA contraction only (WENO5: the contraction noise passes through the non-linear weights)
        assert np.max(np.abs(got - want)) <= (1e-12 if scheme == _lib.ADV_WENO5 else 1e-14) * scale, n
        halo = got.copy()
        og.interior(halo)[...] = 0.5
        assert np.all(halo == 0.5)


def test_argument_checks(cuda):
    grid = ob.RectilinearGrid(size=(4, 4, 4), extent=(1, 1, 1), device=cuda)
    cg = grid.c_grid()
    f = ob.CenterField(grid)
    t = _lib.pointer_table([f.ptr])
    lib = _lib.load()
    assert lib.obm_sinking_tendencies(C.byref(cg), 1, t, t, t, 7, 0, None) == -3       # unknown scheme
    assert lib.obm_sinking_tendencies(C.byref(cg), 99, t, t, t, 0, 0, None) == -2      # too many tracers
    assert lib.obm_sinking_tendencies(C.byref(cg), 1, None, t, t, 0, 0, None) == -1
    assert lib.obm_sinking_tendencies(C.byref(cg), 0, None, None, None, 0, 0, None) == 0


def total_nitrogen(model, sed, grid):
    dz, A = grid.dz.reshape(-1, 1, 1), grid.dx * grid.dy
    names_N = ("NO₃", "NH₄", "P", "Z", "sPOM", "bPOM", "DOM")
    dzt = torch.from_numpy(np.ascontiguousarray(dz)).to(model.tracers["P"].data.device)
    water = sum((model.tracers[n].interior * dzt).sum().item() for n in names_N) * A
    sediment = sum(f.interior.sum().item() for f in sed.fields.values()) * A
    return water + sediment


@pytest.mark.parametrize("advection", ["UpwindBiased1", "UpwindBiased3", "WENO5"])
def test_total_nitrogen_bookkeeping_closes_to_rounding(cuda, advection):
    """Every term on the device — LOBSTER tendencies, sinking of sPOM / bPOM by obm_sinking_tendencies (open bottom),
    SimpleMultiG sediment fed by the same bottom-face flux, forward-Euler tracer update by obm_rk3_substep — with the
    sediment in its Euler mode (AB2, χ = −1/2): water-column N + sediment N is constant to rounding once the
    sediment's one-call lag is closed by a trailing state update."""
    grid = ob.RectilinearGrid(size=(4, 3, 16), extent=(4.0, 3.0, 64.0), device=cuda)
    sed = ob.SimpleMultiGSediment(grid, timestepper="QuasiAdamsBashforth2", chi=-0.5)
    bgc = ob.LOBSTER(grid, oxygen=ob.Oxygen(), sediment=sed, surface_photosynthetically_active_radiation=100.0)
    model = ob.BiogeochemicalModel(grid, bgc, timestepper="Euler", sinking_advection=advection)
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *synthetic.lobster_range(n))
    model.tracers["O₂"].data.fill_(250.0)
    N0 = total_nitrogen(model, sed, grid)
    b0 = model.tracers["bPOM"].interior.sum().item()
    for _ in range(50):
        model.time_step(20.0)
    model.update_state()  # applies the last stored sediment tendency
    N1 = total_nitrogen(model, sed, grid)
    assert abs(N1 - N0) <= 1e-12 * abs(N0), (N0, N1)
    assert sum(f.interior.sum().item() for f in sed.fields.values()) > 0.01  # a visible amount reached the sediment
    assert model.tracers["bPOM"].interior.sum().item() != b0


@pytest.mark.parametrize("sediment_timestepper", ["QuasiAdamsBashforth2", "RungeKutta3"])
def test_reference_sediment_conservation_test(cuda, sediment_timestepper):
    """test/test_sediments.jl:37-80 as written there: rectilinear 3×3×50 grid, 10×10×500 m, sPOM = bPOM = 1, NO₃ = 10,
    NH₄ = 1, O₂ = 1000, 100 RK3 steps of Δt = 1 with the sinking advected on the device, both sediment time steppers;
    total nitrogen to rtol 0.2e-6 and sediment nitrogen non-zero everywhere."""
    grid = ob.RectilinearGrid(size=(3, 3, 50), extent=(10.0, 10.0, 500.0), device=cuda)
    sed = ob.SimpleMultiGSediment(grid, timestepper=sediment_timestepper)
    bgc = ob.LOBSTER(grid, oxygen=ob.Oxygen(), sediment=sed)
    model = ob.BiogeochemicalModel(grid, bgc, timestepper="RungeKutta3", sinking_advection="UpwindBiased3")
    model.set(**{"sPOM": 1.0, "bPOM": 1.0, "NO₃": 10.0, "NH₄": 1.0, "O₂": 1000.0})
    for f in model.tracers.values():
        f.fill_halos_zero_gradient()
    N0 = total_nitrogen(model, sed, grid)
    for _ in range(100):
        model.time_step(1.0)
    N1 = total_nitrogen(model, sed, grid)
    assert abs(N1 - N0) <= 0.2e-6 * abs(N0), (N0, N1)
    assert all(bool((f.interior != 0).all()) for f in sed.fields.values())
