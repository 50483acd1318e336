"""Pin the ORACLE's light kernels against the reference's analytic known answers (CPU only):
test/test_light.jl:10-106 (two-band closed form, multi-band exponentials) and
test/test_PISCES.jl:98-127 (euphotic depth = −10 ln 1000, mixed-layer means)."""
import math

import numpy as np

import oceanbiome_b200 as ob


def column_grid(oracle, Nz, Lz, H=3):
    g = ob.RectilinearGrid(size=(Nz,), extent=(Lz,), topology=("Flat", "Flat", "Bounded"), halo=(H,), device="cpu")
    return g, oracle.Grid.like(g)


def box_grid(oracle, size, extent):
    g = ob.RectilinearGrid(size=size, extent=extent, device="cpu")
    return g, oracle.Grid.like(g)


def test_two_band_closed_form(oracle):
    # test_light.jl:10-50: grid (2,2,2) extent (2,2,2), P = 2.5 + z, PAR⁰ = 100
    g, og = box_grid(oracle, (2, 2, 2), (2, 2, 2))
    par = ob.TwoBandPhotosyntheticallyActiveRadiation(grid=g, surface_PAR=100.0)
    P = np.zeros(og.parent_shape)
    og.interior(P)[...] = (2.5 + g.zc).reshape(-1, 1, 1)
    PAR = oracle.par_twoband(og, par.c_params(), P, 100.0)
    kr, kb = par.water_red_attenuation, par.water_blue_attenuation
    xr, xb = par.chlorophyll_red_attenuation, par.chlorophyll_blue_attenuation
    er, eb, r, R = par.chlorophyll_red_exponent, par.chlorophyll_blue_exponent, par.pigment_ratio, par.phytoplankton_chlorophyll_ratio
    ir = [(2.0 * R / r) ** er * 0.5]
    ib = [(2.0 * R / r) ** eb * 0.5]
    ir.append(ir[0] + (2.0 * R / r) ** er * 0.5 + (1.0 * R / r) ** er * 0.5)
    ib.append(ib[0] + (2.0 * R / r) ** eb * 0.5 + (1.0 * R / r) ** eb * 0.5)
    expected = [100.0 * (math.exp(-0.5 * kr - ir[0] * xr) + math.exp(-0.5 * kb - ib[0] * xb)) / 2,
                100.0 * (math.exp(-1.5 * kr - ir[1] * xr) + math.exp(-1.5 * kb - ib[1] * xb)) / 2]
    got = og.interior(PAR)[:, 0, 0]  # k = 1, 2 (bottom, top)
    np.testing.assert_allclose(got, expected[::-1], rtol=1e-14)
    # SURVEY §8c "readings already validated": identical to the test's closed form
    np.testing.assert_allclose(got[::-1], [87.99032377900511, 70.00421099730072], rtol=1e-14)
    # every column identical
    assert np.all(og.interior(PAR) == got.reshape(-1, 1, 1))


def test_multi_band_exponentials(oracle):
    # test_light.jl:52-106
    g, og = box_grid(oracle, (2, 2, 2), (2, 2, 2))
    P = np.zeros(og.parent_shape)
    og.interior(P)[...] = 2 / 1.31
    m = ob.MultiBandPhotosyntheticallyActiveRadiation(
        grid=g, bands=((1, 2),), base_bands=[1, 2], base_water_attenuation_coefficient=[0.01, 0.01],
        base_chlorophyll_exponent=[2, 2], base_chlorophyll_attenuation_coefficient=[0.1, 0.1], surface_PAR=100.0)
    bands, total = oracle.par_multiband(og, m.c_params(), P, None, 1.31, 100.0)
    expected = 100 * np.exp(g.zc * (0.01 + 0.1 * 2 ** 2))
    np.testing.assert_allclose(og.interior(bands[0])[:, 0, 0], expected, rtol=1e-12)

    m2 = ob.MultiBandPhotosyntheticallyActiveRadiation(
        grid=g, bands=((1, 2), (8, 9)), base_bands=[1, 2, 8, 9],
        base_water_attenuation_coefficient=[0.01, 0.01, 0.02, 0.02], base_chlorophyll_exponent=[2, 2, 1.5, 1.5],
        base_chlorophyll_attenuation_coefficient=[0.1, 0.1, 0.2, 0.2], surface_PAR=100.0)
    bands, total = oracle.par_multiband(og, m2.c_params(), P, None, 1.31, 100.0)
    e1 = 100 * np.exp(g.zc * (0.01 + 0.1 * 2 ** 2)) / 2
    e2 = 100 * np.exp(g.zc * (0.02 + 0.2 * 2 ** 1.5)) / 2
    np.testing.assert_allclose(og.interior(bands[0])[:, 1, 1], e1, atol=1e-4)
    np.testing.assert_allclose(og.interior(bands[1])[:, 1, 1], e2, atol=1e-4)
    np.testing.assert_allclose(og.interior(total)[:, 1, 1], e1 + e2, atol=1e-3)


def test_default_morel_band_means(oracle):
    # default 3-band coefficients from numerical_mean (multi_band.jl:97-104,136-140)
    g, _ = box_grid(oracle, (2, 2, 2), (2, 2, 2))
    m = ob.MultiBandPhotosyntheticallyActiveRadiation(grid=g)
    np.testing.assert_allclose(m.water_attenuation_coefficient, [0.01143, 0.0699895, 0.3737025], rtol=1e-6)
    np.testing.assert_allclose(m.chlorophyll_exponent, [0.673612, 0.649562, 0.65865], rtol=1e-6)
    np.testing.assert_allclose(m.chlorophyll_attenuation_coefficient, [0.09989825, 0.04361825, 0.04255], rtol=1e-6)
    # the oracle's C restatement of numerical_mean agrees with the host mirror
    from oceanbiome_b200.light import MOREL_kʷ, MOREL_λ
    assert abs(oracle.numerical_mean(MOREL_λ, MOREL_kʷ, 400, 500) - m.water_attenuation_coefficient[0]) < 1e-17
    assert m.field_names == ("PAR₁", "PAR₂", "PAR₃")


def _light(z):
    return math.exp(z / 10) if z <= 0 else 2 - math.exp(-z / 10)


def test_euphotic_depth_and_mixed_layer_means(oracle):
    # test_PISCES.jl:95-125: 10 levels over 100 m, PAR = 3 light(z) (incl. the value in the top halo),
    # zₘₓₗ = −25, κ = z > −25 ? 2 : 1
    g, og = column_grid(oracle, 10, 100)
    PAR = np.zeros(og.parent_shape)
    PAR[:, 0, 0] = [3 * _light(z) for z in og.zc_parent]
    zeu = oracle.euphotic_depth(og, PAR)
    assert math.isclose(zeu[0, 0, 0], -10 * math.log(1000), rel_tol=1e-12)
    zmxl = np.full(og.plane_shape, -25.0)
    mean_PAR = oracle.mixed_layer_mean(og, zmxl, PAR)
    assert math.isclose(mean_PAR[0, 0, 0], 3 * 10 / 25 * (1 - math.exp(-25 / 10)), rel_tol=0.1)
    kappa = np.zeros(og.parent_shape)
    kappa[:, 0, 0] = [2.0 if z > -25 else 1.0 for z in og.zc_parent]
    mean_k = oracle.mixed_layer_mean(og, zmxl, kappa)
    assert math.isclose(mean_k[0, 0, 0], 2, rel_tol=0.1)


def test_euphotic_depth_never_reached_returns_znode_zero(oracle):
    # compute_euphotic_depth.jl:28 — bright column: falls back to znode(k = 0), the centre below the bottom
    g, og = column_grid(oracle, 8, 16)
    PAR = np.full(og.parent_shape, 50.0)
    zeu = oracle.euphotic_depth(og, PAR)
    assert zeu[0, 0, 0] == og.zc_parent[og.Hz - 1] == -17.0


# ---- independent restatement (oracle/pyref_light.py: plain Python, one column at a time, no code shared with the C) ----

def _pyref():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import pyref_light
    return pyref_light


def _stretched_column(oracle, rng, Nz, Lz, nx=3):
    faces = -Lz * (1 - np.linspace(0, 1, Nz + 1)) ** 1.7  # faces[0] = −Lz … faces[-1] = 0, fine near the surface
    faces[1:-1] += rng.uniform(-0.2, 0.2, Nz - 1) * np.diff(faces).min()
    g = ob.RectilinearGrid(size=(nx, Nz), x=(0, nx), z=faces, topology=("Periodic", "Flat", "Bounded"), device="cpu")
    return g, oracle.Grid.like(g)


def _cols(ref, g):
    """1-based zc[0 … Nz + 1], zf[1 … Nz + 1] views of the grid's host z arrays."""
    zc = ref.Col(g.zc_host[g.Hz - 1:g.Hz + g.Nz + 1], lo=0)
    zf = ref.Col(g.zf_host[g.Hz:g.Hz + g.Nz + 1], lo=1)
    return zc, zf


def test_c_oracle_matches_independent_two_band(oracle):
    ref, rng = _pyref(), np.random.default_rng(5)
    for Nz, Lz in ((1, 10.0), (7, 80.0), (40, 300.0)):
        g, og = _stretched_column(oracle, rng, Nz, Lz)
        par = ob.TwoBandPhotosyntheticallyActiveRadiation(grid=g, surface_PAR=100.0)
        P = np.zeros(og.parent_shape)
        og.interior(P)[...] = rng.uniform(0.0, 2.0, og.interior(P).shape)
        og.interior(P)[Nz // 2, 0, 1] = 0.0  # a level without phytoplankton
        S = np.zeros(og.plane_shape)
        S[...] = rng.uniform(5.0, 200.0, S.shape)
        PAR = oracle.par_twoband(og, par.c_params(), P, S)
        zc, zf = _cols(ref, g)
        for i in range(g.Nx):
            want = ref.two_band(Nz, zc, zf, ref.Col(og.interior(P)[:, 0, i]), float(og.interior(S)[0, 0, i]),
                                par.water_red_attenuation, par.water_blue_attenuation, par.chlorophyll_red_attenuation,
                                par.chlorophyll_blue_attenuation, par.chlorophyll_red_exponent, par.chlorophyll_blue_exponent,
                                par.pigment_ratio, par.phytoplankton_chlorophyll_ratio)
            np.testing.assert_allclose(og.interior(PAR)[:, 0, i], want.v, rtol=2e-14)


def test_c_oracle_matches_independent_multi_band_and_column_diagnostics(oracle):
    ref, rng = _pyref(), np.random.default_rng(6)
    from oceanbiome_b200.light import MOREL_e, MOREL_kʷ, MOREL_λ, MOREL_χ
    bands = ((400, 500), (500, 600), (600, 700))
    for Nz, Lz in ((2, 20.0), (9, 120.0), (64, 400.0)):
        g, og = _stretched_column(oracle, rng, Nz, Lz)
        m = ob.MultiBandPhotosyntheticallyActiveRadiation(grid=g, surface_PAR=100.0, surface_PAR_division=[0.5, 0.25, 0.25])
        # the three Morel means, derived independently from the base tables
        for mine, base in ((m.water_attenuation_coefficient, MOREL_kʷ), (m.chlorophyll_exponent, MOREL_e),
                           (m.chlorophyll_attenuation_coefficient, MOREL_χ)):
            np.testing.assert_allclose(mine, ref.band_coefficients(bands, list(MOREL_λ), list(base)), rtol=1e-15)
        a, b = np.zeros(og.parent_shape), np.zeros(og.parent_shape)
        og.interior(a)[...] = rng.uniform(0.0, 1.5, og.interior(a).shape)
        og.interior(b)[...] = rng.uniform(0.0, 0.5, og.interior(b).shape)
        PAR0 = 80.0
        fields, total = oracle.par_multiband(og, m.c_params(), a, b, 1.0, PAR0)
        zc, zf = _cols(ref, g)
        zmxl = np.zeros(og.plane_shape)
        zmxl[...] = rng.uniform(-0.9 * Lz, -0.02 * Lz, zmxl.shape)
        # level above the surface as the value boundary condition would leave it (2·surface − top)
        for f, d in zip(fields + [total], m.surface_PAR_division + [1.0]):
            f[og.Hz + Nz] = 2 * PAR0 * d - f[og.Hz + Nz - 1]
        cutoff = 0.3  # reached inside even the shallow columns
        zeu = oracle.euphotic_depth(og, total, cutoff)
        mean = oracle.mixed_layer_mean(og, zmxl, total)
        for i in range(g.Nx):
            chl = ref.Col(og.interior(a)[:, 0, i] + og.interior(b)[:, 0, i])
            want_total = np.zeros(Nz)
            for n in range(3):
                w = ref.multi_band(Nz, zc, chl, PAR0, m.surface_PAR_division[n], m.water_attenuation_coefficient[n],
                                   m.chlorophyll_exponent[n], m.chlorophyll_attenuation_coefficient[n])
                np.testing.assert_allclose(og.interior(fields[n])[:, 0, i], w.v, rtol=2e-14)
                want_total += np.array(w.v)
            np.testing.assert_allclose(og.interior(total)[:, 0, i], want_total, rtol=2e-14)
            col = ref.Col(list(og.interior(total)[:, 0, i]) + [float(total[og.Hz + Nz, og.Hy, og.Hx + i])])
            got_zeu = float(og.interior(zeu)[0, 0, i])
            assert abs(got_zeu - ref.euphotic_depth(Nz, zc, col, cutoff)) <= 1e-12 * Lz
            got_mean = float(og.interior(mean)[0, 0, i])
            want_mean = ref.mixed_layer_mean(Nz, zf, float(og.interior(zmxl)[0, 0, i]), col)
            assert abs(got_mean - want_mean) <= 1e-13 * abs(want_mean)
