"""Host-side mirror of the reference's plugin surface (CPU only, no compute): constructors, tracer
lists, conserved groups and summaries behave like the reference's
(test/test_NutrientsPlanktonDetritus.jl:141-152, docstrings constructors.jl:51-63,163-175)."""
import ctypes as C
import itertools

import numpy as np
import pytest

import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib


@pytest.fixture(scope="module")
def grid():
    return ob.RectilinearGrid(size=(3, 3, 30), extent=(10, 10, 200), device="cpu")


def test_constructor_docstrings(grid):
    # constructors.jl:56-61 and :168-173
    assert repr(ob.LOBSTER(grid)) == ("LOBSTER model (:NO₃, :NH₄, :P, :Z, :sPOM, :bPOM, :DOM) \n"
                                      " Light attenuation: Two-band light attenuation model (Float64)\n"
                                      " Sediment: Nothing\n Particles: Nothing\n Modifiers: Nothing")
    assert repr(ob.NPZD(grid)).startswith("NPZD model (:N, :P, :Z, :T, :D) \n")


def test_constructor_types(grid):
    lobster, npzd = ob.LOBSTER(grid), ob.NPZD(grid)
    assert isinstance(lobster, ob.Biogeochemistry) and isinstance(npzd, ob.Biogeochemistry)
    u = lobster.underlying_biogeochemistry
    assert isinstance(u.nutrients, ob.NitrateAmmonia) and isinstance(u.plankton, ob.PhytoZoo)
    assert isinstance(u.detritus, ob.TwoParticleAndDissolved)
    u = npzd.underlying_biogeochemistry
    assert isinstance(u.nutrients, ob.Nutrient) and isinstance(u.detritus, ob.Detritus)
    assert u.plankton.temperature_coefficient == 1.88
    assert lobster.required_biogeochemical_auxiliary_fields() == ("PAR",)
    assert set(lobster.biogeochemical_auxiliary_fields()) == {"PAR"}


@pytest.mark.parametrize("nut,det,car,oxy", list(itertools.product(
    (ob.Nutrient, ob.NitrateAmmonia, ob.NitrateAmmoniaIron),
    (None, ob.Detritus, ob.TwoParticleAndDissolved, ob.VariableRedfieldDetritus), (0, 1, 3), (False, True))))
def test_tracer_lists_agree_with_the_library(nut, det, car, oxy):
    bgc = ob.NutrientsPlanktonDetritus(nut(), ob.PhytoZoo(), det() if det else None,
                                       ob.CarbonateSystem(car) if car else None, ob.Oxygen() if oxy else None)
    p = bgc.c_params()
    names = ((C.c_char * 16) * _lib.OBM_NPD_MAX_TRACERS)()
    n = _lib.load().obm_npd_tracer_names(C.byref(p), names)
    got = tuple(bytes(names[i]).split(b"\0")[0].decode() for i in range(n))
    assert got == bgc.required_biogeochemical_tracers()


def test_conserved_tracers_and_scalers(grid):
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), scale_negatives=True)
    nitrogen, carbon = bgc.conserved_tracers()
    assert nitrogen == ("P", "Z", "NO₃", "NH₄", "sPOM", "bPOM", "DOM")
    assert carbon["tracers"] == ("P", "Z", "DIC", "sPOM", "bPOM", "DOM")
    R, rho = 6.56, 0.1
    assert carbon["scalefactors"] == ((1 + rho) * R, R, 1, R, R, R)
    assert isinstance(bgc.modifiers, tuple) and len(bgc.modifiers) == 2
    assert all(isinstance(m, ob.ScaleNegativeTracers) for m in bgc.modifiers)
    npzd = ob.NPZD(grid, scale_negatives=True)
    assert isinstance(npzd.modifiers, ob.ScaleNegativeTracers)
    assert npzd.modifiers.tracers == ("P", "Z", "N", "D")
    with pytest.raises(ValueError, match="Incorrect number of scale factors"):
        ob.ScaleNegativeTracers(("A", "B"), scalefactors=(1,))


def test_drift_velocities(grid):
    bgc = ob.LOBSTER(grid)
    assert bgc.biogeochemical_drift_velocity("sPOM") == -3.47e-5
    assert bgc.biogeochemical_drift_velocity("bPOM") == -200 / 86400
    assert bgc.biogeochemical_drift_velocity("NO₃") is None
    assert np.isclose(ob.NPZD(grid).biogeochemical_drift_velocity("D"), -2.7489 / 86400)


def test_grid_layout_matches_oceananigans_parent_arrays():
    g = ob.RectilinearGrid(size=(5, 4, 6), x=(0, 10), y=(0, 8), z=(-12, 0), device="cpu")
    assert g.parent_shape == (12, 10, 11)
    f = ob.CenterField(g)
    assert f.data.stride() == (110, 11, 1)  # x fastest, then y, then z: Julia column-major (x, y, z)
    np.testing.assert_allclose(g.zf, np.linspace(-12, 0, 7))
    np.testing.assert_allclose(g.zc_host, np.arange(-17, 6, 2.0)[:12])
    flat = ob.RectilinearGrid(size=(10,), extent=(100,), topology=("Flat", "Flat", "Bounded"), device="cpu")
    assert (flat.Nx, flat.Ny, flat.Nz, flat.Hx, flat.Hy, flat.Hz) == (1, 1, 10, 0, 0, 3)
    s = ob.RectilinearGrid(size=(8, 16, 4), extent=(8, 16, 4), device="cpu").slab(1, 4)
    assert (s.Nx, s.Ny, s.Nz) == (8, 4, 4) and s.y == (4.0, 8.0)


def test_compute_entry_points_refuse_cpu_tensors(grid):
    bgc = ob.LOBSTER(grid)
    model = ob.BiogeochemicalModel(grid, bgc)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.update_state()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.compute_tendencies()


def test_z_faces_must_increase():
    import pytest
    import oceanbiome_b200 as ob
    with pytest.raises(ValueError, match="strictly increasing"):
        ob.RectilinearGrid(size=(2, 2, 3), x=(0, 1), y=(0, 1), z=[0.0, -1.0, -2.0, -3.0], device="cpu")
    with pytest.raises(ValueError, match="strictly increasing"):
        ob.RectilinearGrid(size=(2, 2, 2), x=(0, 1), y=(0, 1), z=[-2.0, -1.0, -1.0], device="cpu")
    g = ob.RectilinearGrid(size=(2, 2, 3), x=(0, 1), y=(0, 1), z=[-3.0, -2.0, -0.5, 0.0], device="cpu")
    assert list(g.zc) == [-2.5, -1.25, -0.25]


def test_latitude_longitude_grid_geometry():
    import math
    import numpy as np
    import pytest
    import oceanbiome_b200 as ob
    R = 6371e3
    g = ob.LatitudeLongitudeGrid(size=(36, 18, 2), longitude=(-180, 180), latitude=(-90, 90), z=(-2, 0), device="cpu")
    assert (g.Nx, g.Ny, g.Nz) == (36, 18, 2) and g.topology == ("Periodic", "Bounded", "Bounded")
    assert math.isclose(g.cell_area().sum() * g.Nx, 4 * math.pi * R ** 2, rel_tol=1e-14)
    v = g.cell_volume()
    assert tuple(v.shape) == (2, 18, 1) and math.isclose(float(v.sum()) * g.Nx, 4 * math.pi * R ** 2 * 2.0, rel_tol=1e-13)
    assert np.allclose(g.latitude_centers[[0, -1]], [-85.0, 85.0])
    # areas shrink towards the poles and are symmetric about the equator
    a = g.cell_area()
    assert np.allclose(a, a[::-1], rtol=1e-13) and a[0] < a[4] < a[8]
    # slabs keep the class and their own rows; together they are the whole grid
    parts = [g.slab(r, 2) for r in range(2)]
    assert all(isinstance(p, ob.LatitudeLongitudeGrid) for p in parts)
    assert np.allclose(np.concatenate([p.cell_area() for p in parts]), a, rtol=1e-13)
    # the reference's light-test grid (test/test_light.jl:113-114) and argument checking
    t = ob.LatitudeLongitudeGrid(size=(5, 5, 2), longitude=(-180, 180), latitude=(-85, 85), z=(-2, 0), device="cpu")
    assert t.parent_shape == (2 + 6, 5 + 6, 5 + 6) and list(t.zc) == [-1.5, -0.5]
    with pytest.raises(ValueError):
        ob.LatitudeLongitudeGrid(size=(5, 5, 2), longitude=(-180, 180), latitude=(-95, 85), z=(-2, 0), device="cpu")


def test_pisces_latitude_choice_follows_the_reference():
    """PISCES.jl:360-367: a grid with its own latitude overrides a prescribed one (with the reference's warning), a
    RectilinearGrid cannot offer one; ModelLatitude's per-row table holds the latitude and BOTH argument orders of the
    reference's day-length calls (growth_rate.jl:29-30 swapped, :141-143) for every interior row."""
    import pytest
    import oceanbiome_b200 as ob
    g = ob.LatitudeLongitudeGrid(size=(4, 5, 2), longitude=(0, 10), latitude=(-40, 60), z=(-10, 0), device="cpu")
    with pytest.warns(UserWarning, match="prescribed value is ignored"):
        u = ob.PISCES(g, latitude=ob.PrescribedLatitude(45.0)).underlying_biogeochemistry
    assert isinstance(u.latitude, ob.ModelLatitude)
    u = ob.PISCES(g, latitude=ob.ModelLatitude()).underlying_biogeochemistry
    t = 0.3 * 365 * 86400.0
    rows = u.row_table(g, t)
    assert tuple(rows.shape) == (3, 5)
    for j, lat in enumerate(g.latitude_centers):
        assert float(rows[0, j]) == float(lat)
        assert float(rows[1, j]) == u.day_length(float(lat), t) and float(rows[2, j]) == u.day_length(t, float(lat))
    assert float(rows[2, 0]) < float(rows[2, -1])  # late April: the southern rows have the shorter days
    assert u.c_params(t).latitude == 0.0  # the scalar members are not used by the per-row launch
    with pytest.raises(ValueError, match="prescribe a latitude"):
        ob.PISCES(ob.RectilinearGrid(size=(2, 2, 2), extent=(1, 1, 1), device="cpu"), latitude=ob.ModelLatitude())
    # a slab keeps its own rows
    g6 = ob.LatitudeLongitudeGrid(size=(4, 6, 2), longitude=(0, 10), latitude=(-30, 30), z=(-10, 0), device="cpu")
    assert list(g6.slab(1, 2).latitude_centers) == list(g6.latitude_centers[3:])


def test_slab_ranges_by_rate():
    """Rows shared out in proportion to the ranks' host-link rates: contiguous, complete, at least `minimum` each, equal
    rates = equal slabs, and the measured 8-GPU case (two groups 1.43× apart) lands on the expected split."""
    import pytest
    from oceanbiome_b200.distributed import slab_ranges, slab_ranges_by_rate
    assert slab_ranges_by_rate(1024, [3.0] * 8) == slab_ranges(1024, 8)
    r = slab_ranges_by_rate(1024, [1 / 396.0] * 4 + [1 / 277.0] * 4, minimum=8)
    assert r[0][0] == 0 and r[-1][1] == 1024 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    rows = [j1 - j0 for j0, j1 in r]
    assert rows[:4] == [rows[0]] * 4 and rows[4:] == [rows[4]] * 4 and 104 <= rows[0] <= 107 and 149 <= rows[4] <= 152
    # the slower group's stage time with its smaller slab equals the faster group's with its larger one (to a row)
    assert abs(rows[0] * 396.0 - rows[4] * 277.0) <= 396.0
    assert [j1 - j0 for j0, j1 in slab_ranges_by_rate(10, [1, 1, 1])] in ([4, 3, 3], [3, 4, 3], [3, 3, 4])
    assert slab_ranges_by_rate(8, [1, 100]) == [(0, 1), (1, 8)]
    for bad in ((4, [1, 0]), (4, [1, float("nan")]), (1, [1, 1])):
        with pytest.raises(ValueError):
            slab_ranges_by_rate(*bad)
