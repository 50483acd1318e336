"""GPU parity: two-band / multi-band PAR scans, euphotic depth and mixed-layer means against the
oracle and against the reference's analytic answers (test/test_light.jl, test_PISCES.jl:98-127)."""
import math

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import synthetic
from helpers import RTOL_TENDENCY, synthetic_state

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


class M:  # the slice of a model the light hooks read
    def __init__(self, grid, tracers, bgc=None, t=0.0):
        self.grid, self.tracers, self.biogeochemistry, self.clock = grid, tracers, bgc, ob.Clock(t)


@pytest.mark.parametrize("size,topo", [((2, 2, 2), None), ((160, 32), ("Periodic", "Flat", "Bounded")),
                                       ((45, 7, 70), None), ((1, 50, 33), None), ((64,), ("Flat", "Flat", "Bounded"))])
def test_two_band_matches_oracle(cuda, oracle, size, topo):
    kw = dict(topology=topo) if topo else {}
    extent = tuple([10.0 * s for s in size[:-1]] + [137.0])
    if len(size) == 3 and size[2] == 70:  # stretched vertical grid
        grid = ob.RectilinearGrid(size=size, x=(0, 45), y=(0, 7), z=lambda k: -140 * (1 - ((k - 1) / 70) ** 1.7), device=cuda)
    else:
        grid = ob.RectilinearGrid(size=size, extent=extent, device=cuda, **kw)
    dev, host, og = synthetic_state(grid, ["P"], {"P": (0.005, 2.0, True)})
    sdev = ob.Field2D(grid, "sPAR")
    synthetic.fill_torch(sdev, "sPAR", 10.0, 300.0)
    par = ob.TwoBandPhotosyntheticallyActiveRadiation(grid=grid, surface_PAR=sdev)
    par.update_biogeochemical_state(M(grid, dev))
    want = oracle.par_twoband(og, par.c_params(), host["P"], sdev.data.cpu().numpy())
    got = par.field.data.cpu().numpy()
    assert rel(og.interior(got), og.interior(want)) <= RTOL_TENDENCY
    outside = got.copy()
    og.interior(outside)[...] = 0
    assert np.all(outside == 0)  # halos untouched


def test_two_band_reference_closed_form(cuda):
    # test_light.jl:10-50 for the three kinds of surface_PAR (:107-131)
    grid = ob.RectilinearGrid(size=(2, 2, 2), extent=(2, 2, 2), device=cuda)
    for sp, discrete in ((100.0, False), (lambda x, y, t: 100, False), (lambda i, j, grid, clock, fields: 100, True)):
        bgc = ob.NPZD(grid, light_attenuation=ob.TwoBandPhotosyntheticallyActiveRadiation(grid=grid, surface_PAR=sp, discrete_form=discrete))
        model = ob.BiogeochemicalModel(grid, bgc, extra_tracers=("T", "S"))
        model.set(P=torch.tensor(2.5 + grid.zc).reshape(-1, 1, 1))
        model.update_state()
        got = model.auxiliary_fields["PAR"].interior[:, 0, 0].cpu().numpy()
        np.testing.assert_allclose(got[::-1], [87.99032377900511, 70.00421099730072], rtol=1e-13)


def test_two_band_on_the_reference_latitude_longitude_grid(cuda):
    # test_light.jl:113-122: the same closed form on LatitudeLongitudeGrid(size = (5, 5, 2), …, z = (-2, 0)) — the scan
    # sees sizes, halos and z only, so every column carries the rectilinear answer
    grid = ob.LatitudeLongitudeGrid(size=(5, 5, 2), longitude=(-180, 180), latitude=(-85, 85), z=(-2, 0), device=cuda)
    bgc = ob.NPZD(grid, light_attenuation=ob.TwoBandPhotosyntheticallyActiveRadiation(grid=grid, surface_PAR=100.0))
    model = ob.BiogeochemicalModel(grid, bgc, extra_tracers=("T", "S"))
    model.set(P=torch.tensor(2.5 + grid.zc).reshape(-1, 1, 1))
    model.update_state()
    got = model.auxiliary_fields["PAR"].interior.cpu().numpy()
    np.testing.assert_allclose(got[::-1, 0, 0], [87.99032377900511, 70.00421099730072], rtol=1e-13)
    assert np.all(got == got[:, :1, :1])


@pytest.mark.parametrize("nbands", [1, 2, 3, 4])
def test_multi_band_matches_oracle(cuda, oracle, nbands):
    grid = ob.RectilinearGrid(size=(70, 3, 45), extent=(70, 3, 400), device=cuda)
    dev, host, og = synthetic_state(grid, ["PChl", "DChl"], {"PChl": (0.01, 1.5, True), "DChl": (0.01, 1.5, True)})
    bands = [(400, 500), (500, 600), (600, 700), (350, 400)][:nbands]
    m = ob.MultiBandPhotosyntheticallyActiveRadiation(grid=grid, bands=bands, surface_PAR=83.0)

    class B:
        def chlorophyll(self, model):
            return model.tracers["PChl"], model.tracers["DChl"], 1.0
    m.update_biogeochemical_state(M(grid, dev, B()))
    bo, to = oracle.par_multiband(og, m.c_params(), host["PChl"], host["DChl"], 1.0, 83.0)
    for n, name in enumerate(m.field_names):
        assert rel(og.interior(m.fields[name].data.cpu().numpy()), og.interior(bo[n])) <= RTOL_TENDENCY
    assert rel(og.interior(m.total.data.cpu().numpy()), og.interior(to)) <= RTOL_TENDENCY


def test_multi_band_reference_exponentials(cuda):
    # test_light.jl:52-106 through NPZD (Chl = 1.31 P)
    grid = ob.RectilinearGrid(size=(2, 2, 2), extent=(2, 2, 2), device=cuda)
    m2 = ob.MultiBandPhotosyntheticallyActiveRadiation(
        grid=grid, bands=((1, 2), (8, 9)), base_bands=[1, 2, 8, 9], base_water_attenuation_coefficient=[0.01, 0.01, 0.02, 0.02],
        base_chlorophyll_exponent=[2, 2, 1.5, 1.5], base_chlorophyll_attenuation_coefficient=[0.1, 0.1, 0.2, 0.2], surface_PAR=100.0)
    model = ob.BiogeochemicalModel(grid, ob.NPZD(grid, light_attenuation=m2))
    model.set(P=2 / 1.31)
    model.update_state()
    aux = model.auxiliary_fields
    assert list(aux) == ["PAR", "PAR₁", "PAR₂"]
    e1 = 100 * np.exp(grid.zc * (0.01 + 0.1 * 2 ** 2)) / 2
    e2 = 100 * np.exp(grid.zc * (0.02 + 0.2 * 2 ** 1.5)) / 2
    np.testing.assert_allclose(aux["PAR₁"].interior[:, 0, 0].cpu().numpy(), e1, atol=1e-4)
    np.testing.assert_allclose(aux["PAR₂"].interior[:, 1, 1].cpu().numpy(), e2, atol=1e-4)
    np.testing.assert_allclose(aux["PAR"].interior[:, 1, 0].cpu().numpy(), e1 + e2, atol=1e-3)
    model.time_step(1.0)  # "check all the models work as expected"


def test_euphotic_depth_and_mixed_layer_means(cuda, oracle):
    # analytic (test_PISCES.jl:98-127) …
    grid = ob.RectilinearGrid(size=(10,), extent=(100,), topology=("Flat", "Flat", "Bounded"), device=cuda)
    light = lambda z: math.exp(z / 10) if z <= 0 else 2 - math.exp(-z / 10)  # noqa: E731
    PAR = ob.CenterField(grid, "PAR")
    PAR.data[:, 0, 0] = torch.tensor([3 * light(z) for z in grid.zc_host], dtype=torch.float64)
    zeu, mean = ob.Field2D(grid), ob.Field2D(grid)
    ob.compute_euphotic_depth(zeu, PAR)
    assert math.isclose(zeu.data.item(), -10 * math.log(1000), rel_tol=1e-12)
    zmxl = ob.Field2D(grid, fill=-25.0)
    ob.compute_mixed_layer_mean(mean, zmxl, PAR, grid)
    assert math.isclose(mean.data.item(), 3 * 10 / 25 * (1 - math.exp(-25 / 10)), rel_tol=0.1)
    ob.compute_mixed_layer_mean(mean, zmxl, 2.0, grid)
    assert mean.data.item() == 2.0
    # … and against the oracle on a ragged 3-D grid with random mixed-layer depths and a bright patch
    grid = ob.RectilinearGrid(size=(53, 9, 40), extent=(53, 9, 400), device=cuda)
    dev, host, og = synthetic_state(grid, ["chl"], {"chl": (0.01, 3.0, True)})
    m = ob.MultiBandPhotosyntheticallyActiveRadiation(grid=grid, surface_PAR=120.0)

    class B:
        def chlorophyll(self, model):
            return model.tracers["chl"], None, 1.0
    m.update_biogeochemical_state(M(grid, dev, B()))
    m.total.interior[:, :2, :5] = 40.0  # columns that never reach the cutoff → znode(k = 0)
    zeu, mean, zmxl = ob.Field2D(grid), ob.Field2D(grid), ob.Field2D(grid)
    synthetic.fill_torch(zmxl, "zmxl", -150.0, -10.0)
    ob.compute_euphotic_depth(zeu, m.total)
    ob.compute_mixed_layer_mean(mean, zmxl, m.total, grid)
    PARh = m.total.data.cpu().numpy()
    zo = oracle.euphotic_depth(og, PARh)
    mo = oracle.mixed_layer_mean(og, zmxl.data.cpu().numpy(), PARh)
    assert rel(og.interior(zeu.data.cpu().numpy()), og.interior(zo)) <= RTOL_TENDENCY
    assert rel(og.interior(mean.data.cpu().numpy()), og.interior(mo)) <= RTOL_TENDENCY
    assert og.interior(zo)[0, 0, 0] == og.zc_parent[og.Hz - 1]


@pytest.mark.parametrize("Nz", [1, 3, 5, 31, 35, 97])
def test_scans_at_depths_that_do_not_fill_a_lane_group(cuda, oracle, Nz):
    """The scans give every lane four consecutive levels of a 32-level z-tile (r04): level counts that leave the last
    group of four, the last tile, or both partly empty — and a single level — for the two-band sum scan, the N-band
    product scan and the fused column diagnostics, on 37 × 3 columns (the column tile is ragged too)."""
    grid = ob.RectilinearGrid(size=(37, 3, Nz), x=(0, 37), y=(0, 3), z=lambda k: -6.0 * Nz * (1 - ((k - 1) / Nz) ** 1.3), device=cuda)
    dev, host, og = synthetic_state(grid, ["P", "PChl", "DChl"], {"P": (0.005, 2.0, True), "PChl": (0.001, 1.0, True), "DChl": (0.001, 1.0, True)})
    sdev = ob.Field2D(grid, "sPAR")
    synthetic.fill_torch(sdev, "sPAR", 10.0, 300.0)
    par = ob.TwoBandPhotosyntheticallyActiveRadiation(grid=grid, surface_PAR=sdev)
    par.update_biogeochemical_state(M(grid, dev))
    want = oracle.par_twoband(og, par.c_params(), host["P"], sdev.data.cpu().numpy())
    assert rel(og.interior(par.field.data.cpu().numpy()), og.interior(want)) <= RTOL_TENDENCY
    m = ob.MultiBandPhotosyntheticallyActiveRadiation(grid=grid, surface_PAR=83.0)

    class B:
        def chlorophyll(self, model):
            return model.tracers["PChl"], model.tracers["DChl"], 1.0
    zmxl, zeu, mean = ob.Field2D(grid), ob.Field2D(grid), ob.Field2D(grid)
    synthetic.fill_torch(zmxl, "zmxl", -5.0 * Nz, -0.5)
    for column_state in (None, (zmxl, 1 / 1000, zeu, mean)):
        for f in m.fields.values():
            f.data.zero_()
        m.update_biogeochemical_state(M(grid, dev, B()), column_state=column_state)
        bo, to = oracle.par_multiband(og, m.c_params(), host["PChl"], host["DChl"], 1.0, 83.0)
        for n, name in enumerate(m.field_names):
            assert rel(og.interior(m.fields[name].data.cpu().numpy()), og.interior(bo[n])) <= RTOL_TENDENCY
        assert rel(og.interior(m.total.data.cpu().numpy()), og.interior(to)) <= RTOL_TENDENCY
    PARh = m.total.data.cpu().numpy()
    zo, mo = oracle.euphotic_depth(og, PARh), oracle.mixed_layer_mean(og, zmxl.data.cpu().numpy(), PARh)
    assert rel(og.interior(zeu.data.cpu().numpy()), og.interior(zo)) <= RTOL_TENDENCY
    assert rel(og.interior(mean.data.cpu().numpy()), og.interior(mo)) <= RTOL_TENDENCY
