"""CPU: the sinking-advection oracle (oracle_sinking.c).  Oceananigans' operator is not in the reference tree, so the
restatement is pinned by properties: flux-form telescoping, exact face values for polynomials of the scheme's degree,
the upwind direction, and untouched halos / sub-range behaviour."""
import numpy as np
import pytest

UP1, CEN2, UP3 = 0, 1, 2


def make(oracle, Nz=12, stretched=True, seed=0):
    rng = np.random.default_rng(seed)
    Hz = 3
    if stretched:
        dz = 1.0 + 0.5 * np.sin(np.arange(Nz + 2 * Hz) * 0.7) ** 2
    else:
        dz = np.full(Nz + 2 * Hz, 2.0)
    zf = np.concatenate([[0.0], np.cumsum(dz)])
    zf -= zf[Hz + Nz]  # surface at 0
    zc = 0.5 * (zf[:-1] + zf[1:])
    g = oracle.Grid(4, 3, Nz, 2, 1, Hz, zc, zf)
    return g, rng, zf, zc


def w_field(g, value):
    """z-face field: `value` on faces 0 … Nz−1, 0 on the closed top face (setup_velocity_fields, open bottom)."""
    w = np.zeros(g.parent_shape)
    w[g.Hz:g.Hz + g.Nz] = value
    return w


@pytest.mark.parametrize("scheme", [UP1, CEN2, UP3])
@pytest.mark.parametrize("wval", [-3e-3, 2e-3])
def test_flux_form_telescopes(oracle, scheme, wval):
    g, rng, zf, _ = make(oracle)
    c = rng.uniform(0.1, 2.0, size=g.parent_shape)
    w = w_field(g, wval) * rng.uniform(0.5, 1.5, size=g.parent_shape)  # face-varying speed, one sign
    G = [np.full(g.parent_shape, 7.0)]
    oracle.sinking_tendencies(g, [c], [w], G, scheme, accumulate=False)
    dz = np.diff(zf)[g.Hz:g.Hz + g.Nz].reshape(-1, 1, 1)
    column_change = (g.interior(G[0]) * dz).sum(axis=0)
    # only the bottom face carries a flux (top face: w = 0); its face value is first order in every scheme but CEN2
    wb = w[g.Hz, g.Hy:g.Hy + g.Ny, g.Hx:g.Hx + g.Nx]
    below, above = c[g.Hz - 1, g.Hy:g.Hy + g.Ny, g.Hx:g.Hx + g.Nx], c[g.Hz, g.Hy:g.Hy + g.Ny, g.Hx:g.Hx + g.Nx]
    face = (below + above) / 2 if scheme == CEN2 else (below if wval > 0 else above)
    np.testing.assert_allclose(column_change, wb * face, rtol=1e-12)
    halo = G[0].copy()
    g.interior(halo)[...] = 7.0
    assert np.all(halo == 7.0)


def test_accumulate_adds(oracle):
    g, rng, _, _ = make(oracle, seed=3)
    c = rng.uniform(0.1, 2.0, size=g.parent_shape)
    w = w_field(g, -1e-3)
    A = [np.zeros(g.parent_shape)]
    oracle.sinking_tendencies(g, [c], [w], A, UP1, accumulate=False)
    B = [np.full(g.parent_shape, 0.25)]
    oracle.sinking_tendencies(g, [c], [w], B, UP1, accumulate=True)
    assert np.array_equal(g.interior(B[0]), 0.25 + g.interior(A[0]))


def test_upwind_direction_and_uniform_field(oracle):
    g, rng, zf, _ = make(oracle, stretched=False)
    # a uniform tracer sinking at uniform speed: interior cells do not change, the bottom cell neither (in = out),
    # the top cell loses w c / Δz (nothing comes in through the closed surface)
    c = np.full(g.parent_shape, 1.5)
    w = w_field(g, -4e-3)
    for scheme in (UP1, CEN2, UP3):
        G = [np.zeros(g.parent_shape)]
        oracle.sinking_tendencies(g, [c], [w], G, scheme, accumulate=False)
        Gi = g.interior(G[0])
        assert np.all(Gi[:-1] == 0.0)
        np.testing.assert_allclose(Gi[-1], -4e-3 * 1.5 / 2.0, rtol=1e-15)
    # a single spike moves DOWN under w < 0 with first-order upwind: the cell below gains exactly what the spike loses
    c = np.zeros(g.parent_shape)
    c[g.Hz + 6] = 1.0
    G = [np.zeros(g.parent_shape)]
    oracle.sinking_tendencies(g, [c], [w], G, UP1, accumulate=False)
    Gi = g.interior(G[0])
    assert np.all(Gi[6] < 0) and np.all(Gi[5] == -Gi[6]) and np.all(Gi[7] == 0)


def test_third_order_face_values_are_exact_for_quadratics(oracle):
    """UpwindBiased(order=3) reconstructs the face value of a quadratic's CELL AVERAGES exactly on a uniform grid, so
    away from the boundaries the tendency equals the exact −w (c(z_{k+1}) − c(z_k)) / Δz."""
    g, rng, zf, zc = make(oracle, Nz=16, stretched=False)
    dz = 2.0
    q = lambda z: 0.3 + 0.05 * z + 0.002 * z * z  # noqa: E731
    Q = lambda z: 0.3 * z + 0.025 * z ** 2 + 0.002 / 3 * z ** 3  # noqa: E731  antiderivative
    cbar = (Q(zf[1:]) - Q(zf[:-1])) / dz
    c = np.broadcast_to(cbar.reshape(-1, 1, 1), g.parent_shape).copy()
    for wval in (-2e-3, 2e-3):
        w = np.full(g.parent_shape, wval)  # open top and bottom: every face carries a flux
        G = [np.zeros(g.parent_shape)]
        oracle.sinking_tendencies(g, [c], [w], G, UP3, accumulate=False)
        zfi = zf[g.Hz:g.Hz + g.Nz + 1]
        exact = -wval * (q(zfi[1:]) - q(zfi[:-1])) / dz
        got = g.interior(G[0])[:, 0, 0]
        np.testing.assert_allclose(got[3:-3], exact[3:-3], rtol=1e-11)
