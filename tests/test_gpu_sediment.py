"""GPU: fused sediment kernels against the oracle (same assumptions, see oracle_sediment.c), integer bottom indices
bit-exact, and the one property the reference intended to test (test/test_sediments.jl:37-80): total nitrogen —
water column ∫dV plus sediment ∫dA — is conserved while particles sink into the sediment."""
import math

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import synthetic
from helpers import synthetic_state

pytestmark = pytest.mark.gpu


def host(f):
    return np.ascontiguousarray(f.data.cpu().numpy())


def lobster_with(cuda, sediment_ctor, **kw):
    grid = ob.RectilinearGrid(size=(24, 6, 12), extent=(240.0, 60.0, 120.0), device=cuda)
    sed = sediment_ctor(grid, **kw)
    bgc = ob.LOBSTER(grid, oxygen=ob.Oxygen(), sediment=sed, surface_photosynthetically_active_radiation=100.0)
    model = ob.BiogeochemicalModel(grid, bgc)
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *synthetic.lobster_range(n))
    return grid, sed, bgc, model


@pytest.mark.parametrize("ctor,kw", [
    (ob.SimpleMultiGSediment, dict(timestepper="QuasiAdamsBashforth2")),
    (ob.SimpleMultiGSediment, dict(timestepper="RungeKutta3", advection="Centered2")),
    (ob.InstantRemineralisationSediment, dict(sinking_tracers=("sPOM", "bPOM"), remineralisation_reciever="NH₄")),
])
def test_hooks_match_oracle(cuda, oracle, ctor, kw):
    grid, sed, bgc, model = lobster_with(cuda, ctor, **kw)
    og = oracle.Grid.like(grid)
    for n, f in sed.fields.items():
        synthetic.fill_torch(f, "sed" + n, 0.01, 10.0, log=True)
    for d in (sed.Gn, sed.Gm):
        for n, f in d.items():
            synthetic.fill_torch(f, "G" + n + str(id(d) % 7), -1e-6, 1e-6)
    b = sed.biogeochemistry
    h_tr = {n: host(model.tracers[n]) for n in model.tracers}
    h_w = {n: host(sed._w_field(bgc, n)) for n in b.sinking_fluxes()}
    h_pools = [host(f) for f in sed.fields.values()]
    h_Gn = [host(f) for f in sed.Gn.values()]
    h_Gm = [host(f) for f in sed.Gm.values()]
    h_tracked = [np.zeros(og.plane_shape) for _ in sed.tracked_fields]
    for n in b.coupled_tracers():
        synthetic.fill_torch(model.Gn[n], "Gc" + n, -1e-7, 1e-7)
    h_Gc = [host(model.Gn[n]) for n in b.coupled_tracers()]
    fo = oracle.sediment_fields(NO3=h_tr.get("NO₃") if b.required_tracers() else None, NH4=h_tr.get("NH₄") if b.required_tracers() else None,
                                O2=h_tr.get("O₂") if b.required_tracers() else None,
                                sinking=[h_tr[n] for n in b.sinking_fluxes()], sinking_w=[h_w[n] for n in b.sinking_fluxes()],
                                pools=h_pools, Gn=h_Gn, Gm=h_Gm, tracked=h_tracked, G_coupled=h_Gc)
    p = sed.c_params()
    dt = 120.0
    model.clock.last_stage_dt = dt
    sed.last_dt = dt  # not the first step: AB2 with χ = 0.1
    sed.update_biogeochemical_state(model)
    sed.update_tendencies(bgc, model)
    oracle.sediment_update_state(og, p, fo, dt, chi=0.1)
    oracle.sediment_update_tendencies(og, p, fo)
    close = lambda a, b_: np.testing.assert_allclose(a, b_, rtol=1e-12, atol=1e-300)  # noqa: E731
    for f, h in zip(sed.fields.values(), h_pools):
        close(host(f), h)
    for f, h in zip(sed.Gn.values(), h_Gn):
        close(host(f), h)
    for f, h in zip(sed.Gm.values(), h_Gm):
        if sed.timestepper == "RungeKutta3":
            close(host(f), h)  # the tendency of the sediment's second stage, recomputed inside the hook
        else:
            assert np.array_equal(host(f), h)  # cached tendency: a copy
    for f, h in zip(sed.tracked_fields.values(), h_tracked):
        close(host(f), h)
    for n, h in zip(b.coupled_tracers(), h_Gc):
        got = host(model.Gn[n])
        scale = np.abs(h).max()
        assert np.max(np.abs(got - h)) <= 1e-12 * scale, n
        # only the bottom plane was touched
        assert np.array_equal(got[og.Hz + 1:], host(model.Gn[n])[og.Hz + 1:])


def test_bottom_indices_bit_exact(cuda, oracle):
    grid = ob.RectilinearGrid(size=(37, 9, 20), extent=(37, 9, 200), device=cuda)
    og = oracle.Grid.like(grid)
    h = ob.Field2D(grid, "bottom_height")
    synthetic.fill_torch(h, "bottom_height", -230.0, 15.0)
    got = ob.calculate_bottom_indices(grid, h).cpu().numpy()
    want = oracle.find_bottom_cells(og, host(h))
    assert np.array_equal(og.interior(got), og.interior(want))
    assert got.dtype == np.int64 and og.interior(got).min() == 1 and og.interior(got).max() == 20


def test_total_nitrogen_is_conserved_with_sinking_into_the_sediment(cuda):
    """Column model: BGC tendencies + first-order upwind sinking of sPOM/bPOM (closed top, open bottom into the
    sediment) + SimpleMultiG.  Σ N·V (water) + Σ (Ns + Nf + Nr)·A (sediment) stays constant (rtol 2e-7 in the
    reference's commented-out test; forward Euler here)."""
    grid = ob.RectilinearGrid(size=(4, 3, 16), extent=(4.0, 3.0, 64.0), device=cuda)
    # AB2 with χ = −1/2 is a forward-Euler pool update: every stored tendency is applied exactly once with weight Δt,
    # so the bookkeeping closes to rounding (χ = 0.1 leaves a (½+χ)·Δt·G imbalance at the end of the run, the
    # sediment's own RK3 an O(Δt²) one — the reference's test allows 2e-7 for that, see test_gpu_sinking.py)
    sed = ob.SimpleMultiGSediment(grid, timestepper="QuasiAdamsBashforth2", chi=-0.5)
    bgc = ob.LOBSTER(grid, oxygen=ob.Oxygen(), sediment=sed, surface_photosynthetically_active_radiation=100.0)
    model = ob.BiogeochemicalModel(grid, bgc, timestepper="Euler")
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *synthetic.lobster_range(n))
    model.tracers["O₂"].data.fill_(250.0)
    dz, A = grid.dz[0], grid.dx * grid.dy
    names_N = ("NO₃", "NH₄", "P", "Z", "sPOM", "bPOM", "DOM")

    def total():
        water = sum(model.tracers[n].interior.sum().item() for n in names_N) * dz * A
        sediment = sum(f.interior.sum().item() for f in sed.fields.values()) * A
        return water + sediment

    def sinking_fluxes():  # upwind-1: downward flux through the lower face of every cell, from the CURRENT state
        return {n: -bgc.biogeochemical_drift_velocity(n) * model.tracers[n].interior.clone() for n in ("sPOM", "bPOM")}

    def sink(fluxes, dt):  # flux divergence; nothing enters through the surface, the bottom face feeds the sediment
        for n, flux in fluxes.items():
            c = model.tracers[n].interior
            c -= dt * flux / dz
            c[:-1] += dt * flux[1:] / dz

    N0 = total()
    dt = 20.0
    for _ in range(50):
        # one step: state update (sediment pools stepped with the PREVIOUS flux-based Gⁿ), tendencies, tracers
        model.update_state()          # the sediment tracks the bottom flux of THIS state …
        fluxes = sinking_fluxes()     # … and the water column loses exactly the same flux
        model.compute_tendencies()
        for n, c in model.tracers.items():
            c.data.add_(model.Gn[n].data, alpha=dt)
        sink(fluxes, dt)
        model.clock.last_stage_dt = dt
    # one more state update applies the last stored sediment tendency so both sides have seen 50 fluxes … the
    # sediment lags the water column by exactly one step (it integrates the flux computed at the previous call)
    model.update_state()
    N1 = total()
    assert abs(N1 - N0) <= 1e-12 * abs(N0), (N0, N1)
    assert sum(f.interior.sum().item() for f in sed.fields.values()) > 0.01  # a visible amount reached the sediment
    assert all(bool((f.interior > 0).all()) for f in sed.fields.values())


@pytest.mark.parametrize("timestepper,with_sediment", [("RungeKutta3", True), ("Euler", False)])
def test_column_ensemble_run_as_a_replayed_graph_is_bit_identical_to_the_eager_loop(cuda, timestepper, with_sediment):
    """SURVEY §8 f-3, the 1-D half: an ensemble of independent columns (LOBSTER + carbonates + O₂, sinking POM, optionally the
    SimpleMultiG sediment with its own stepper) stepped on the device.  `run(graph=True)` — first step eager, then one
    captured time step replayed — leaves every tracer, every G⁻, the sediment pools and every snapshot exactly as the
    eager loop of `time_step` does, and the clock where it belongs; PISCES (host-evaluated day lengths) is refused."""
    def build():
        grid = ob.RectilinearGrid(size=(48, 2, 20), extent=(48.0, 2.0, 200.0), device=cuda)
        sed = ob.SimpleMultiGSediment(grid) if with_sediment else None
        bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen(), sediment=sed, scale_negatives=True,
                         surface_photosynthetically_active_radiation=100.0)
        model = ob.BiogeochemicalModel(grid, bgc, timestepper=timestepper, sinking_advection="UpwindBiased3")
        for n, f in model.tracers.items():
            synthetic.fill_torch(f, n, *synthetic.lobster_range(n))
        if sed is not None:
            for n, f in sed.fields.items():
                synthetic.fill_torch(f, "sed" + n, 1e-2, 10.0, True)
        return model, sed

    steps, every, dt = 9, 3, 120.0
    (a, sa), (b, sb) = build(), build()
    ra = a.run(dt, steps, graph=False, output_every=every)
    rb = b.run(dt, steps, graph=True, output_every=every)
    torch.cuda.synchronize()
    for n in a.tracers:
        assert torch.equal(a.tracers[n].data, b.tracers[n].data), n
        if a.Gm is not None:
            assert torch.equal(a.Gm[n].data, b.Gm[n].data), n
    for n in ra:
        assert torch.equal(ra[n], rb[n]) and ra[n].shape == (steps // every, 20, 2, 48), n
    if with_sediment:
        for n in sa.fields:
            assert torch.equal(sa.fields[n].data, sb.fields[n].data), n
        assert sa.iteration == sb.iteration and sa.last_dt == sb.last_dt
    assert a.clock.iteration == b.clock.iteration == steps and abs(a.clock.time - b.clock.time) < 1e-9
    assert a.clock.last_stage_dt == b.clock.last_stage_dt
    assert not torch.equal(b.tracers["P"].data, build()[0].tracers["P"].data)  # it did move
    # and it continues: three more steps either way
    a.run(dt, 3)
    b.run(dt, 3, graph=True)
    torch.cuda.synchronize()
    for n in a.tracers:
        assert torch.equal(a.tracers[n].data, b.tracers[n].data), n
    grid = ob.RectilinearGrid(size=(4, 2, 6), extent=(4.0, 2.0, 60.0), device=cuda)
    with pytest.raises(ValueError, match="not PISCES"):
        ob.BiogeochemicalModel(grid, ob.PISCES(grid)).run(60.0, 5, graph=True)
