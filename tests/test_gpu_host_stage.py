"""The host-buffer (end-to-end) path: HostStagedStage pipelines H2D → kernels → D2H over x–y slabs and must give
bit-identical tendencies to the device-resident path (slabs are independent: every kernel is pointwise or
column-local).  Also checks partial launches through `grid.restrict`."""
import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import pisces, synthetic
from oceanbiome_b200.host_stage import HostStagedStage

pytestmark = pytest.mark.gpu


def lobster(cuda, size=(48, 20, 12)):
    grid = ob.RectilinearGrid(size=size, extent=(480.0, 200.0, 120.0), device=cuda)
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen(), scale_negatives=True,
                     surface_photosynthetically_active_radiation=90.0)
    model = ob.BiogeochemicalModel(grid, bgc)
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *synthetic.lobster_range(n))
    model.tracers["NO₃"].interior[2, 3, :7] = -0.5  # exercise the scaling
    return grid, bgc, model


@pytest.mark.parametrize("engine,Nx", [("sm", 48), ("sm", 47), ("dma", 48)])  # Nx = 47: rows are not 16-byte multiples
@pytest.mark.parametrize("nslabs", [1, 3, 20])
def test_host_staged_lobster_equals_resident(cuda, nslabs, engine, Nx):
    grid, bgc, model = lobster(cuda, size=(Nx, 20, 12))
    stage = HostStagedStage(model, nslabs=nslabs, copy_engine=engine)
    stage.upload_from_device()
    before = {n: f.data.clone() for n, f in model.tracers.items()}
    # resident reference
    model.update_state()
    for g in model.Gn.values():
        g.data.zero_()
    bgc.underlying_biogeochemistry.compute_tendencies(grid, model.tracers, bgc.biogeochemical_auxiliary_fields(), model.Gn, accumulate=False)
    want = {n: model.Gn[n].data.clone() for n in stage.gnames}
    # scramble the device state, then run host → host
    for n, f in model.tracers.items():
        f.data.fill_(float("nan"))
    for g in model.Gn.values():
        g.data.fill_(-7.0)
    stage.step()
    stage.synchronize()
    for n in stage.gnames:
        got = grid.interior(stage.host_G[n])
        assert torch.equal(got, grid.interior(want[n]).cpu()), n
    assert stage.h2d_bytes == len(model.tracers) * grid.Ny * (grid.Nx + 6) * grid.Nz * 8  # interior k-planes only


@pytest.mark.parametrize("engine", ["dma", "sm"])
def test_host_staged_stage_can_return_the_rescaled_tracers(cuda, engine):
    """`return_tracers=True`: what `ScaleNegativeTracers` did to the state in place comes back to the host arrays too."""
    grid, bgc, model = lobster(cuda)
    stage = HostStagedStage(model, nslabs=3, copy_engine=engine, return_tracers=True)
    stage.upload_from_device()
    sent = {n: t.clone() for n, t in stage.host_tracers.items()}
    model.update_state()  # resident reference: the state update rescales in place
    want = {n: f.data.clone() for n, f in model.tracers.items()}
    for f in model.tracers.values():
        f.data.fill_(float("nan"))
    stage.step()
    stage.synchronize()
    for n in stage.names:
        assert torch.equal(grid.interior(stage.host_tracers[n]), grid.interior(want[n]).cpu()), n
    assert bool((grid.interior(sent["NO₃"]) < 0).any()) and bool((grid.interior(stage.host_tracers["NO₃"]) >= 0).all())
    assert not torch.equal(grid.interior(stage.host_tracers["NH₄"]), grid.interior(sent["NH₄"]))  # its group was rescaled
    per_field = grid.Ny * (grid.Nx + 6) * grid.Nz * 8
    assert stage.d2h_bytes == (len(stage.gnames) + len(stage.names)) * per_field


def test_host_staged_pisces_equals_resident(cuda):
    grid = ob.RectilinearGrid(size=(40, 16, 10), extent=(4e3, 1.6e3, 200.0), device=cuda)
    bgc = ob.PISCES(grid, scale_negatives=True, surface_photosynthetically_active_radiation=75.0)
    model = ob.BiogeochemicalModel(grid, bgc)
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *pisces.synthetic_range(n))
    pisces.fill_synthetic_auxiliary(bgc, model)
    stage = HostStagedStage(model, nslabs=4)
    stage.upload_from_device()
    model.update_state()
    bgc.underlying_biogeochemistry.compute_tendencies(grid, model.tracers, bgc.biogeochemical_auxiliary_fields(), model.Gn, accumulate=False)
    want = {n: model.Gn[n].data.clone() for n in stage.gnames}
    for g in model.Gn.values():
        g.data.fill_(-7.0)
    stage.step()
    stage.synchronize()
    for n in stage.gnames:
        assert torch.equal(grid.interior(stage.host_G[n]), grid.interior(want[n]).cpu()), n


def test_partial_launch_touches_only_the_subrange(cuda):
    grid, bgc, model = lobster(cuda)
    model.update_state()
    G = {n: ob.CenterField(grid, fill=5.0) for n in model.tracers}
    with grid.restrict(4, 9):
        bgc.underlying_biogeochemistry.compute_tendencies(grid, model.tracers, bgc.biogeochemical_auxiliary_fields(), G, accumulate=False)
    P = G["P"].interior
    assert bool((P[:, :4] == 5.0).all()) and bool((P[:, 9:] == 5.0).all()) and bool((P[:, 4:9] != 5.0).all())
