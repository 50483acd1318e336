"""CPU: the sinking-advection oracle (oracle_sinking.c).  Oceananigans' operator is not in the reference tree, so the
restatement is pinned by properties: flux-form telescoping, exact face values for polynomials of the scheme's degree,
the upwind direction, and untouched halos / sub-range behaviour."""
import numpy as np
import pytest

UP1, CEN2, UP3, WENO5 = 0, 1, 2, 3


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


@pytest.mark.parametrize("scheme", [UP1, CEN2, UP3, WENO5])
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
    for scheme in (UP1, CEN2, UP3, WENO5):
        G = [np.zeros(g.parent_shape)]
        oracle.sinking_tendencies(g, [c], [w], G, scheme, accumulate=False)
        Gi = g.interior(G[0])
        if scheme == WENO5:  # the normalised weights return c·Σα/Σα: c to an ulp, not to the bit
            assert np.all(np.abs(Gi[:-1]) <= 4e-3 * 1.5 / 2.0 * 1e-15)
        else:
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
        zfi = zf[g.Hz:g.Hz + g.Nz + 1]
        exact = -wval * (q(zfi[1:]) - q(zfi[:-1])) / dz
        for scheme in (UP3, WENO5):  # every WENO candidate stencil is exact for a quadratic, whatever the weights
            G = [np.zeros(g.parent_shape)]
            oracle.sinking_tendencies(g, [c], [w], G, scheme, accumulate=False)
            got = g.interior(G[0])[:, 0, 0]
            np.testing.assert_allclose(got[4:-4], exact[4:-4], rtol=1e-11)


def _weno5_z(um2, um1, u0, up1, up2, eps=1e-8):
    """Second transcription of WENO5 with Z weights from the published formulas (Jiang & Shu 1996 eq. 2.15-2.17 for the
    candidates and β_k; Borges et al. 2008 eq. 25-28 for τ₅ and α_k) — written without looking at oracle_sinking.c."""
    q = [(2 * um2 - 7 * um1 + 11 * u0) / 6, (-um1 + 5 * u0 + 2 * up1) / 6, (2 * u0 + 5 * up1 - up2) / 6]
    beta = [13 / 12 * (um2 - 2 * um1 + u0) ** 2 + 1 / 4 * (um2 - 4 * um1 + 3 * u0) ** 2,
            13 / 12 * (um1 - 2 * u0 + up1) ** 2 + 1 / 4 * (um1 - up1) ** 2,
            13 / 12 * (u0 - 2 * up1 + up2) ** 2 + 1 / 4 * (3 * u0 - 4 * up1 + up2) ** 2]
    tau5 = abs(beta[0] - beta[2])
    alpha = [d * (1 + (tau5 / (b + eps)) ** 2) for d, b in zip((1 / 10, 6 / 10, 3 / 10), beta)]
    return sum(a * v for a, v in zip(alpha, q)) / sum(alpha)


@pytest.mark.parametrize("wval", [-2e-3, 2e-3])
def test_weno5_interior_faces_equal_an_independent_transcription(oracle, wval):
    """Rough data (uniform random + a step): the interior tendency −(F_{k+1} − F_k)/Δz with F = w·c̃ and c̃ from the second
    transcription above, upwind side chosen by the sign of w; WENO3 / first order only within two cells of the ends."""
    g, rng, zf, _ = make(oracle, Nz=20, stretched=False, seed=8)
    col = rng.uniform(0.1, 2.0, g.Nz + 2 * g.Hz)
    col[g.Hz + 9:] += 3.0                                                 # a step
    c = np.broadcast_to(col.reshape(-1, 1, 1), g.parent_shape).copy()
    w = np.full(g.parent_shape, wval)
    G = [np.zeros(g.parent_shape)]
    oracle.sinking_tendencies(g, [c], [w], G, WENO5, accumulate=False)
    u = col[g.Hz:g.Hz + g.Nz]                                             # interior cells 0 … Nz−1

    def face(k):                                                          # face k between cells k−1 and k
        if wval > 0:
            return _weno5_z(u[k - 3], u[k - 2], u[k - 1], u[k], u[k + 1])
        return _weno5_z(u[k + 2], u[k + 1], u[k], u[k - 1], u[k - 2])
    got = g.interior(G[0])[:, 0, 0]
    for k in range(4, g.Nz - 4):
        want = -wval * (face(k + 1) - face(k)) / 2.0
        assert abs(got[k] - want) <= 1e-13 * max(abs(want), abs(wval)), k


def test_weno5_does_not_create_new_extrema_at_a_step(oracle):
    """Essentially non-oscillatory: advecting a step with WENO5 over-/undershoots by far less than the third-order linear
    scheme does (forward Euler, CFL 0.2, 40 steps)."""
    g, rng, zf, _ = make(oracle, Nz=40, stretched=False)
    col = np.where(np.arange(g.Nz + 2 * g.Hz) > g.Hz + 20, 1.0, 0.0)
    w = np.full(g.parent_shape, -1.0)
    out = {}
    for scheme in (UP3, WENO5):
        c = np.broadcast_to(col.reshape(-1, 1, 1), g.parent_shape).copy()
        for _ in range(40):
            G = [np.zeros(g.parent_shape)]
            oracle.sinking_tendencies(g, [c], [w], G, scheme, accumulate=False)
            g.interior(c)[...] += 0.4 * g.interior(G[0])                 # Δt = 0.4: CFL = 0.2 at Δz = 2
        out[scheme] = g.interior(c)[5:-5, 0, 0]
    over = {s: max(v.max() - 1.0, -v.min()) for s, v in out.items()}
    assert over[UP3] > 1e-2 and over[WENO5] < 0.1 * over[UP3], over
