"""GPU parity: the fused NPD tendency kernel (through the C ABI) against the oracle on identical
seeded inputs, for every model variant the reference tests (test_NutrientsPlanktonDetritus.jl:41-96)
plus the BASELINE.json configs C1 (NPZD 160×1×32) and C2 (LOBSTER+carbonate+O₂ columns).
Tolerance: 1e-12 scale-aware relative (helpers.RTOL_TENDENCY)."""
import itertools

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import synthetic
from helpers import RTOL_TENDENCY, assert_tendency_parity, synthetic_state

pytestmark = pytest.mark.gpu

NUTRIENTS = (ob.NitrateAmmonia, ob.NitrateAmmoniaIron, ob.Nutrient)
DETRITUS = (None, ob.Detritus, ob.TwoParticleAndDissolved, ob.VariableRedfieldDetritus)


def run_both(oracle, grid, bgc, PAR_range=(0.0, 150.0), accumulate=False, g0=0.25):
    names = bgc.required_biogeochemical_tracers()
    dev, host, og = synthetic_state(grid, names, synthetic.lobster_range)
    pdev, phost, _ = synthetic_state(grid, ["PAR"], {"PAR": (*PAR_range, False)})
    G = {n: ob.CenterField(grid, "G" + n, fill=g0 if accumulate else 7.0) for n in names}
    bgc.compute_tendencies(grid, dev, {"PAR": pdev["PAR"]}, G, accumulate=accumulate)
    torch.cuda.synchronize()
    Go = oracle.npd_tendencies(og, bgc.c_params(), [host[n] for n in names], phost["PAR"],
                               G=[np.full(og.parent_shape, g0) for _ in names] if accumulate else None,
                               accumulate=accumulate)
    got = {n: og.interior(G[n].data.cpu().numpy()) for n in names}
    want = {n: og.interior(g) for n, g in zip(names, Go)}
    So = oracle.npd_tendency_scales(og, bgc.c_params(), [host[n] for n in names], phost["PAR"])
    scales = {n: og.interior(s) for n, s in zip(names, So)}
    return names, got, (want, scales), G, og


def assert_parity(names, got, want_scales, offset=0.0, label="npd"):
    """|Δ| ≤ 1e-12·max(|want|, S) per tendency, S = Σ|additive terms| of that tendency (oracle.npd_tendency_scales);
    pure relative error asserted where the tendency is not a near-total cancellation (helpers.assert_tendency_parity)."""
    want, S = want_scales
    return assert_tendency_parity(label, [n for n in names if n != "T"], got, want, S, offset)[0]


@pytest.mark.parametrize("nut,det,car,oxy", list(itertools.product(NUTRIENTS, DETRITUS, (0, 1), (False, True))))
def test_every_variant_matches_oracle(cuda, oracle, nut, det, car, oxy):
    grid = ob.RectilinearGrid(size=(37, 5, 9), extent=(37, 5, 90), device=cuda)  # ragged: not a multiple of anything
    bgc = ob.NutrientsPlanktonDetritus(nut(), ob.PhytoZoo(), det() if det else None,
                                       ob.CarbonateSystem(car) if car else None, ob.Oxygen() if oxy else None)
    names, got, want, G, og = run_both(oracle, grid, bgc)
    assert_parity(names, got, want, label=f"npd[{nut.__name__},{det.__name__ if det else None},{car},{oxy}]")
    # halos of G are never written
    for n in names:
        full = G[n].data.cpu().numpy().copy()
        og.interior(full)[...] = 7.0
        assert np.all(full == 7.0)


def test_npzd_readme_grid_C1(cuda, oracle):
    # BASELINE.json configs[0]: NPZD + TwoBandPAR on the README grid 160×1×32 (Flat y), halo (3, 3)
    grid = ob.RectilinearGrid(size=(160, 32), extent=(10e3, 500), topology=("Periodic", "Flat", "Bounded"), device=cuda)
    bgc = ob.NPZD(grid).underlying_biogeochemistry
    names = bgc.required_biogeochemical_tracers()
    dev, host, og = synthetic_state(grid, names, synthetic.RANGES_NPZD)
    pdev, phost, _ = synthetic_state(grid, ["PAR"], {"PAR": (0.0, 100.0, False)})
    G = {n: ob.CenterField(grid) for n in names}
    bgc.compute_tendencies(grid, dev, pdev, G, accumulate=False)
    Go = oracle.npd_tendencies(og, bgc.c_params(), [host[n] for n in names], phost["PAR"])
    So = oracle.npd_tendency_scales(og, bgc.c_params(), [host[n] for n in names], phost["PAR"])
    got = {n: og.interior(G[n].data.cpu().numpy()) for n in names}
    want = {n: og.interior(g) for n, g in zip(names, Go)}
    assert_parity(names, got, (want, {n: og.interior(x) for n, x in zip(names, So)}), label="npzd_c1")
    assert np.all(got["T"] == 0.0)


def test_lobster_columns_C2_accumulate(cuda, oracle):
    # BASELINE.json configs[1]: LOBSTER + carbonates + O₂, 4096 columns × 64 levels; the
    # update_tendencies! seam adds into an existing Gⁿ (accumulate = 1)
    grid = ob.RectilinearGrid(size=(4096, 64), extent=(4096.0, 200.0), topology=("Periodic", "Flat", "Bounded"),
                              device=cuda)
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen()).underlying_biogeochemistry
    names, got, want, _, _ = run_both(oracle, grid, bgc, accumulate=True, g0=1e-7)
    assert_parity(names, got, want, offset=1e-7, label="lobster_c2_accumulate")


def test_carbonate_replicates_share_one_tendency(cuda, oracle):
    # CarbonateSystem(3): DIC1..3 / Alk1..3 all receive the same tendency (carbonate_system.jl:70-83)
    grid = ob.RectilinearGrid(size=(16, 4, 8), extent=(16, 4, 80), device=cuda)
    bgc = ob.NutrientsPlanktonDetritus(ob.NitrateAmmonia(), ob.PhytoZoo(), ob.TwoParticleAndDissolved(),
                                       ob.CarbonateSystem(3), None)
    names, got, want, _, _ = run_both(oracle, grid, bgc)
    assert_parity(names, got, want, label="npd_carbonate_replicates")
    assert np.array_equal(got["DIC1"], got["DIC3"]) and np.array_equal(got["Alk1"], got["Alk2"])


def test_zero_state_stays_exactly_zero(cuda):
    # test_NutrientsPlanktonDetritus.jl:101-112 through the model driver (RK3, Δt = 1)
    grid = ob.RectilinearGrid(size=(1, 1, 1), extent=(1, 1, 2), device=cuda)
    for nut, det in itertools.product(NUTRIENTS, DETRITUS[1:]):
        bgc = ob.Biogeochemistry(ob.NutrientsPlanktonDetritus(nut(), ob.PhytoZoo(), det(), ob.CarbonateSystem(), ob.Oxygen()),
                                 light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(ob.CenterField(grid, "PAR", 100.0)))
        model = ob.BiogeochemicalModel(grid, bgc)
        model.time_step(1.0)
        assert all(bool((f.interior == 0).all()) for f in model.tracers.values())


def test_conservation_over_100_steps_on_device(cuda):
    # test_NutrientsPlanktonDetritus.jl:114-139: ΣN and ΣC (with scale factors) unchanged after 100 steps
    grid = ob.RectilinearGrid(size=(8, 4, 4), extent=(8, 4, 8), device=cuda)
    bgc = ob.Biogeochemistry(
        ob.NutrientsPlanktonDetritus(ob.NitrateAmmonia(), ob.PhytoZoo(), ob.TwoParticleAndDissolved(), ob.CarbonateSystem(), ob.Oxygen()),
        light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(ob.CenterField(grid, "PAR", 100.0)))
    model = ob.BiogeochemicalModel(grid, bgc)
    for n, f in model.tracers.items():
        lo, hi, log = synthetic.lobster_range(n)
        synthetic.fill_torch(f, n, lo, hi, log)
    cons = bgc.conserved_tracers(labeled=True)

    def total():
        N = sum(model.tracers[n].interior.sum().item() for n in cons["nitrogen"])
        Cc = sum(model.tracers[n].interior.sum().item() * f for n, f in zip(cons["carbon"]["tracers"], cons["carbon"]["scalefactors"]))
        return N, Cc

    N0, C0 = total()
    for _ in range(100):
        model.time_step(1.0)
    N1, C1 = total()
    assert abs(N1 - N0) <= 1.5e-8 * abs(N0) and abs(C1 - C0) <= 1.5e-8 * abs(C0)
    assert model.clock.iteration == 100 and abs(model.clock.time - 100.0) < 1e-9


def test_per_tracer_call_form(cuda, oracle):
    """`bgc(Val(name), …)` — one tracer's tendency at a state — is the fused kernel evaluated on boxes."""
    bgc = ob.LOBSTER(ob.BoxModelGrid(1, device=cuda), carbonate_system=ob.CarbonateSystem(),
                     oxygen=ob.Oxygen()).underlying_biogeochemistry
    state = {"NO₃": 10.0, "NH₄": 0.1, "P": 0.1, "Z": 0.01, "sPOM": 0.2, "bPOM": 0.1, "DOM": 0.3, "DIC": 2200.0,
             "Alk": 2400.0, "O₂": 250.0}
    names = bgc.required_biogeochemical_tracers()
    # the oracle on the same single cell
    grid = ob.BoxModelGrid(1, device=cuda)
    _, host, og = synthetic_state(grid, names, synthetic.lobster_range)
    for n in names:
        og.interior(host[n])[...] = state.get(n, 0.0)
    PARh = np.zeros(og.parent_shape)
    og.interior(PARh)[...] = 40.0
    Go = oracle.npd_tendencies(og, bgc.c_params(), [host[n] for n in names], PARh)
    want = {n: og.interior(g).item() for n, g in zip(names, Go)}
    S = max(abs(v) for v in want.values())
    for n in names:
        got = bgc(n, PAR=40.0, device=cuda, **state)
        assert isinstance(got, float) and abs(got - want[n]) <= RTOL_TENDENCY * S, (n, got, want[n])
    # arrays of states side by side
    PAR = torch.linspace(0.0, 100.0, 7, dtype=torch.float64)
    gP = bgc("P", PAR=PAR, device=cuda, **state)
    assert gP.shape == (7,) and bool((gP[1:] > gP[:-1]).all())  # growth increases with light
    with pytest.raises(KeyError):
        bgc("Q", PAR=1.0, device=cuda)
    with pytest.raises(KeyError):
        bgc("P", PAR=1.0, device=cuda, Q=1.0)
