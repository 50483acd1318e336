"""Pin the ORACLE's ScaleNegativeTracers / ZeroNegativeTracers on the exact outcomes the reference
tests assert: test/test_utils.jl:7-40 and test/test_PISCES.jl:129-168 (CPU only)."""
import math

import numpy as np

import oceanbiome_b200 as ob

PISCES_GROUPS = [  # PISCES/coupling_utils.jl:11-33 in the applied order carbon, iron, phosphate, silicon, nitrogen
    (("P", "D", "Z", "M", "DOC", "POC", "GOC", "DIC", "CaCO₃"), (1,) * 9),
    (("PFe", "DFe", "Z", "M", "SFe", "BFe", "Fe"), (1, 1, 10e-6, 10e-6 * 1.5, 1, 1, 1)),  # placeholder iron ratios, fixed below
    (("P", "D", "Z", "M", "DOC", "POC", "GOC", "PO₄"), (1 / 122,) * 7 + (1,)),
    (("DSi", "Si", "PSi"), (1, 1, 1)),
    (("NH₄", "NO₃", "P", "D", "Z", "M", "DOC", "POC", "GOC"), (1, 1) + (16 / 122,) * 7),
]


def one_cell(oracle):
    g = ob.RectilinearGrid(size=(1, 1, 1), extent=(1, 1, 1), device="cpu")
    return g, oracle.Grid.like(g)


def cell(og, v):
    a = np.zeros(og.parent_shape)
    og.interior(a)[...] = v
    return a


def test_npzd_scaling(oracle):
    # test_utils.jl:7-22: N = 2, P = −1 → N ≈ 1, P ≈ 0 (group (P, Z, N, D), unit scale factors)
    g, og = one_cell(oracle)
    names = ("P", "Z", "N", "D")
    f = [cell(og, v) for v in (-1.0, 0.0, 2.0, 0.0)]
    oracle.scale_negative_tracers(og, f, oracle.make_groups(names, [(names, (1, 1, 1, 1))]))
    vals = [og.interior(a)[0, 0, 0] for a in f]
    assert vals == [0.0, 0.0, 1.0, 0.0]


def test_zeroing(oracle):
    # test_utils.jl:24-40: N = 2, P = −1, Z = −1 excluded → 2, 0, −1
    g, og = one_cell(oracle)
    N, P = cell(og, 2.0), cell(og, -1.0)
    oracle.zero_negative_tracers([N, P])
    assert og.interior(N)[0, 0, 0] == 2.0 and og.interior(P)[0, 0, 0] == 0.0
    nan = cell(og, float("nan"))
    oracle.zero_negative_tracers([nan])
    assert math.isnan(og.interior(nan)[0, 0, 0])  # Julia max propagates NaN


def test_pisces_negativity_protection(oracle):
    # test_PISCES.jl:147-167 with the default iron ratios 0.01 (micro) / 0.015 (meso) of zooplankton/defaults.jl
    g, og = one_cell(oracle)
    groups = list(PISCES_GROUPS)
    groups[1] = (groups[1][0], (1, 1, 0.01, 0.015, 1, 1, 1))
    names = ("P", "D", "Z", "M", "PChl", "DChl", "PFe", "DFe", "DSi", "DOC", "POC", "GOC", "SFe", "BFe", "PSi", "NO₃",
             "NH₄", "PO₄", "Fe", "Si", "CaCO₃", "DIC", "Alk", "O₂")
    state = {n: 0.0 for n in names}
    state.update({"P": -1, "D": 1, "Z": 1, "M": 1, "DOC": 1, "POC": 1, "GOC": 1, "DIC": 1, "CaCO₃": 1, "PO₄": 1})
    f = {n: cell(og, v) for n, v in state.items()}
    cg = oracle.make_groups(names, groups)
    oracle.scale_negative_tracers(og, [f[n] for n in names], cg)
    val = lambda n: og.interior(f[n])[0, 0, 0]  # noqa: E731
    assert val("P") == 0
    for n in ("D", "Z", "M", "DOC", "POC", "GOC", "DIC", "CaCO₃"):
        assert math.isclose(val(n), 7 / 8, rel_tol=1e-8)
    assert val("PO₄") == 1
    # Si = −1, DSi = 0.1: total silicon negative → invalid fill (NaN)
    og.interior(f["Si"])[...] = -1
    og.interior(f["DSi"])[...] = 0.1
    oracle.scale_negative_tracers(og, [f[n] for n in names], cg)
    assert math.isnan(val("DSi"))
    # Fe = −1, Z = 1000, M = 0 → Fe = 0, Z ≈ 900
    og.interior(f["Fe"])[...] = -1
    og.interior(f["Z"])[...] = 1000
    og.interior(f["M"])[...] = 0
    oracle.scale_negative_tracers(og, [f[n] for n in names], cg)
    assert val("Fe") == 0
    assert math.isclose(val("Z"), 900, rel_tol=1e-8)


def test_group_order_matters(oracle):
    """Overlapping groups are applied sequentially (OceanBioME.jl:169), so permuting them changes the result."""
    g, og = one_cell(oracle)
    names = ("A", "B", "C", "D")
    g1, g2 = (("A", "B", "C"), (1, 1, 1)), (("C", "D"), (1, 1))
    r = []
    for order in ([g1, g2], [g2, g1]):
        f = [cell(og, v) for v in (-1.0, 1.0, 2.0, -1.0)]
        oracle.scale_negative_tracers(og, f, oracle.make_groups(names, order))
        r.append([og.interior(a)[0, 0, 0] for a in f])
    np.testing.assert_allclose(r[0], [0.0, 2 / 3, 1 / 3, 0.0], rtol=1e-15)  # g1: B, C ×2/3; g2: C = 4/3·(1/3)/(4/3)
    np.testing.assert_allclose(r[1], [0.0, 0.5, 0.5, 0.0], rtol=1e-15)      # g2: C = 1; g1: B, C ×1/2


def test_immersed_cells_are_left_alone(oracle):
    """`if !immersed_cell(i, j, k, grid)` (negative_tracers.jl:194,253) with a grid-fitted bottom: cells below the
    bottom-most active cell of a column (1-based `bottom_indices`, bottom_indices.jl:19-26) keep their values, negative or
    not; the active cells of the same column are rescaled as usual."""
    g = ob.RectilinearGrid(size=(3, 2, 5), extent=(3, 2, 50), device="cpu")
    bottom = np.ones(oracle.Grid.like(g).plane_shape, dtype=np.int64)
    og0 = oracle.Grid.like(g)
    og0.interior(bottom)[0, :, :] = [[1, 3, 5], [2, 1, 4]]   # bottom-most active cell (1-based k) of the six columns
    og = oracle.Grid(g.Nx, g.Ny, g.Nz, g.Hx, g.Hy, g.Hz, g.zc_host, g.zf_host, bottom)
    names = ("N", "P")
    N, P = cell(og, 2.0), cell(og, -1.0)
    oracle.scale_negative_tracers(og, [N, P], oracle.make_groups(names, [(names, (1, 1))]))
    for j in range(2):
        for i in range(3):
            kb = og.interior(bottom)[0, j, i] - 1
            for k in range(5):
                want = (2.0, -1.0) if k < kb else (1.0, 0.0)
                assert (og.interior(N)[k, j, i], og.interior(P)[k, j, i]) == want, (i, j, k)
