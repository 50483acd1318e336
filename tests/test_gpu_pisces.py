"""GPU parity: the fused PISCES tendency kernel (C ABI) against the oracle on identical seeded inputs,
the reference's conservation / zero-state test through the device path, and the whole PISCES state update
(3-band PAR → zₑᵤ → mixed-layer means → Ω) against the oracle.  Tolerance 1e-12 scale-aware."""
import math

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import pisces, synthetic
from helpers import RTOL_CARBON, RTOL_TENDENCY, assert_tendency_parity, scale_aware_error

pytestmark = pytest.mark.gpu


def build(cuda, size, extent, **kw):
    grid = ob.RectilinearGrid(size=size, extent=extent, device=cuda)
    bgc = ob.PISCES(grid, surface_photosynthetically_active_radiation=120.0, **kw)
    model = ob.BiogeochemicalModel(grid, bgc)
    return grid, bgc, model


def host_aux(og, grid, bgc):
    a = bgc.biogeochemical_auxiliary_fields()
    h = lambda f: np.ascontiguousarray(f.data.cpu().numpy())  # noqa: E731
    return {"PAR1": h(a["PAR₁"]), "PAR2": h(a["PAR₂"]), "PAR3": h(a["PAR₃"]), "PAR": h(a["PAR"]), "Omega": h(a["Ω"]),
            "wPOC": h(a["wPOC"]), "wGOC": h(a["wGOC"]), "mixed_layer_depth_xy": h(a["zₘₓₗ"]),
            "euphotic_depth_xy": h(a["zₑᵤ"]), "mean_mixed_layer_vertical_diffusivity_xy": h(a["κ"]),
            "mean_mixed_layer_light_xy": h(a["mixed_layer_PAR"])}


def fill(model, bgc):
    host = {}
    for n, f in model.tracers.items():
        lo, hi, log = pisces.synthetic_range(n)
        synthetic.fill_torch(f, n, lo, hi, log)
        host[n] = np.ascontiguousarray(f.data.cpu().numpy())
    pisces.fill_synthetic_auxiliary(bgc, model)
    return host


def compare(oracle, og, u, host, aux, G, t, accumulate=False, g0=0.0, label="pisces"):
    """CUDA vs oracle with the stated per-tendency metric: |Δ| ≤ 1e-12·max(|want|, S), S = Σ|additive terms| of that
    tendency (oracle.pisces_tendency_scales); the pure relative error is asserted too (helpers.assert_tendency_parity).
    Returns ({tracer: scale-aware error}, {tracer: pure relative max})."""
    from helpers import tendency_parity
    params = u.c_params(t)
    tr = [host[n] for n in pisces.TRACERS]
    Go = oracle.pisces_tendencies(og, params, tr, aux,
                                  G=[np.full(og.parent_shape, g0) if n < 24 else None for n in range(26)] if accumulate else None,
                                  accumulate=accumulate)
    So = oracle.pisces_tendency_scales(og, params, tr, aux)
    want = {n: og.interior(g) for n, g in zip(pisces.TRACERS, Go) if g is not None}
    S = {n: og.interior(s) for n, s in zip(pisces.TRACERS, So) if s is not None}
    got = {n: og.interior(G[n].data.cpu().numpy()) for n in want}
    assert_tendency_parity(label, list(want), got, want, S, offset=g0 if accumulate else 0.0)
    rows = {n: tendency_parity(got[n], want[n], S[n], g0 if accumulate else 0.0) for n in want}
    return {n: r[0] for n, r in rows.items()}, {n: r[1] for n, r in rows.items()}


@pytest.mark.parametrize("size,t", [((33, 7, 19), 0.37 * 365 * 86400.0), ((128, 2, 40), 1.6)])
def test_fused_tendencies_match_oracle(cuda, oracle, size, t):
    grid, bgc, model = build(cuda, size, (1e4, 1e3, 400.0))
    u = bgc.underlying_biogeochemistry
    host = fill(model, bgc)
    model.clock.time = t
    model.update_state()  # PAR bands, zeu, ML means, Ω on the device
    og = oracle.Grid.like(grid)
    aux = host_aux(og, grid, bgc)
    G = {n: ob.CenterField(grid, fill=3.0) for n in pisces.TRACERS}
    u.compute_tendencies(grid, model.tracers, bgc.biogeochemical_auxiliary_fields(), G, accumulate=False, time=t)
    worst, rel = compare(oracle, og, u, host, aux, G, t, label=f"pisces{size}")
    assert max(worst.values()) <= RTOL_TENDENCY, worst
    # T and S receive nothing; halos untouched
    for n in ("T", "S"):
        assert bool((G[n].data == 3.0).all())
    full = G["P"].data.cpu().numpy().copy()
    og.interior(full)[...] = 3.0
    assert np.all(full == 3.0)


@pytest.mark.parametrize("accumulate", [False, True])
def test_nan_inputs_propagate_like_the_reference(cuda, oracle, accumulate):
    """NaN in an input that the fast path's select-based min / max could swallow (O₂, Ω, zₘₓₗ, zₑᵤ, κ̄: the
    `needs_exact` guard) or that makes a result non-finite (a tracer, PAR): every tendency of such a cell comes from
    the exact pass, NaN exactly where the reference's NaN-propagating `min` / `max` put them, and — accumulate mode —
    nothing is added twice.  All other cells are untouched by the detour."""
    grid, bgc, model = build(cuda, (24, 6, 10), (1e4, 1e3, 300.0))
    u = bgc.underlying_biogeochemistry
    fill(model, bgc)
    model.update_state()
    aux_dev = bgc.biogeochemical_auxiliary_fields()
    nan = float("nan")
    model.tracers["O₂"].interior[3, 1, 2] = nan
    aux_dev["Ω"].interior[4, 2, 5] = nan
    aux_dev["zₘₓₗ"].interior[0, 3, 7] = nan          # a whole column
    aux_dev["zₑᵤ"].interior[0, 4, 9] = nan
    aux_dev["κ"].interior[0, 5, 11] = nan
    model.tracers["P"].interior[6, 0, 13] = nan
    model.tracers["Fe"].interior[7, 1, 15] = float("inf")
    aux_dev["PAR₂"].interior[8, 2, 17] = nan
    model.tracers["T"].interior[2, 3, 19] = nan
    og = oracle.Grid.like(grid)
    host = {n: np.ascontiguousarray(f.data.cpu().numpy()) for n, f in model.tracers.items()}
    aux = host_aux(og, grid, bgc)
    g0 = 1e-7 if accumulate else 3.0
    G = {n: ob.CenterField(grid, fill=g0) for n in pisces.TRACERS}
    u.compute_tendencies(grid, model.tracers, aux_dev, G, accumulate=accumulate, time=0.0)
    torch.cuda.synchronize()
    Go = oracle.pisces_tendencies(og, u.c_params(0.0), [host[n] for n in pisces.TRACERS], aux,
                                  G=[np.full(og.parent_shape, g0) if n < 24 else None for n in range(26)] if accumulate else None,
                                  accumulate=accumulate)
    want = {n: og.interior(g) for n, g in zip(pisces.TRACERS, Go) if g is not None}
    got = {n: og.interior(G[n].data.cpu().numpy()) for n in want}
    off = g0 if accumulate else 0.0
    So = oracle.pisces_tendency_scales(og, u.c_params(0.0), [host[n] for n in pisces.TRACERS], aux)
    Sn = {n: og.interior(s) + off for n, s in zip(pisces.TRACERS, So) if s is not None}
    poisoned = 0
    for n in want:
        bad_w, bad_g = ~np.isfinite(want[n]), ~np.isfinite(got[n])
        assert np.array_equal(np.isnan(want[n]), np.isnan(got[n])), n
        assert np.array_equal(bad_w, bad_g) and np.array_equal(want[n][bad_w & ~np.isnan(want[n])], got[n][bad_g & ~np.isnan(got[n])]), n
        ok = ~bad_w & np.isfinite(Sn[n])
        den = np.maximum(np.abs(want[n][ok]), Sn[n][ok])
        assert np.max(np.abs(got[n][ok] - want[n][ok]) / np.where(den == 0, 1.0, den)) <= RTOL_TENDENCY, n
        poisoned += int(bad_w.sum())
    assert poisoned >= 9 * 3  # every injected NaN reaches several tendencies


def test_extreme_finite_states_match_the_reference_patterns(cuda, oracle):
    """Finite but extreme inputs (zeros, 1e±30, positive biomasses far below their pigment incl. denormals — the states a
    blown-up run hands over just before its first NaN): NaNs are then BORN mid-way (Inf − Inf of overflowed quotas, 0·Inf),
    not read.  Wherever the reference's NaN-propagating arithmetic (the oracle) gives NaN / ±Inf the kernel gives the
    same, every other tendency meets the stated metric — i.e. the fast pass's selects swallowed nothing and the exact
    pass redid every such cell (r01 recorded the denormal-biomass case as a known hole; closed by the operand order of
    min(L_N, L_PO₄, L_Fe, L_Si), tests/test_oracle_pisces.py::test_quota_overflow_reaches_the_exact_pass)."""
    n = 4096
    grid, bgc, model = build(cuda, (n, 1, 1), (1e5, 10.0, 40.0))
    u = bgc.underlying_biogeochemistry
    fill(model, bgc)
    model.update_state()
    rng = np.random.default_rng(20261018)
    aux_dev = bgc.biogeochemical_auxiliary_fields()
    picks = np.array([0.0, 1e-30, 1e-6, 1.0, 1e6, 1e30])
    tiny = np.array([0.0, 5e-324, 1e-310, 1e-280, 1e-200, 1e-30, 1e-6, 1.0, 1e6])
    for name in pisces.TRACERS[:24]:
        src = tiny if name in ("P", "D", "Z", "M") else picks
        v = rng.choice(src, n)
        keep = rng.random(n) < 0.5  # half of the entries stay ordinary, so that single extremes are seen in isolation too
        cur = model.tracers[name].interior[0, 0, :].cpu().numpy()
        model.tracers[name].interior[0, 0, :] = torch.from_numpy(np.where(keep, cur, v)).to(cuda)
    og = oracle.Grid.like(grid)
    host = {m: np.ascontiguousarray(f.data.cpu().numpy()) for m, f in model.tracers.items()}
    aux = host_aux(og, grid, bgc)
    G = {m: ob.CenterField(grid, fill=3.0) for m in pisces.TRACERS}
    u.compute_tendencies(grid, model.tracers, aux_dev, G, accumulate=False, time=0.0)
    torch.cuda.synchronize()
    params = u.c_params(0.0)
    tr = [host[m] for m in pisces.TRACERS]
    Go = oracle.pisces_tendencies(og, params, tr, aux)
    # Fe′ = (−Δ + √(Δ² + 4K·Fe)) / 2K is cancellation noise in these states (Fe = 1e30 against ligands of 0.6, or DOC = 1e6
    # against Fe = 0.1: ten and more digits lost, so one ulp in exp — CUDA's vs libm's — decides the value): the three iron
    # tendencies are measured against Σ|terms| with Fe′'s own Σ|terms| counted (oracle_pisces.c::free_iron)
    oracle.set_nested_free_iron_scale(True)
    try:
        So = oracle.pisces_tendency_scales(og, params, tr, aux)
    finally:
        oracle.set_nested_free_iron_scale(False)
    nonfinite = 0
    for m, g, sc in zip(pisces.TRACERS[:24], Go, So):
        want, got, S = og.interior(g), og.interior(G[m].data.cpu().numpy()), og.interior(sc)
        assert np.array_equal(np.isnan(want), np.isnan(got)), m
        inf = np.isinf(want)
        assert np.array_equal(inf, np.isinf(got)) and np.array_equal(want[inf], got[inf]), m
        ok = np.isfinite(want) & np.isfinite(S)
        den = np.maximum(np.maximum(np.abs(want[ok]), S[ok]), 1e-200)  # below that the products are subnormal: FMA vs mul + add differ
        assert np.max(np.abs(got[ok] - want[ok]) / den) <= RTOL_TENDENCY, m
        nonfinite += int((~np.isfinite(want)).sum())
    assert nonfinite > 100  # the states do produce mid-way NaNs / overflows


def test_accumulate_into_existing_tendencies(cuda, oracle):
    grid, bgc, model = build(cuda, (20, 5, 12), (1e4, 1e3, 300.0))
    u = bgc.underlying_biogeochemistry
    host = fill(model, bgc)
    model.update_state()
    og = oracle.Grid.like(grid)
    aux = host_aux(og, grid, bgc)
    G = {n: ob.CenterField(grid, fill=1e-7) for n in pisces.TRACERS}
    u.compute_tendencies(grid, model.tracers, bgc.biogeochemical_auxiliary_fields(), G, accumulate=True, time=0.0)
    worst, _ = compare(oracle, og, u, host, aux, G, 0.0, accumulate=True, g0=1e-7, label="pisces_accumulate")
    assert max(worst.values()) <= RTOL_TENDENCY, worst


def test_reference_conservation_test_on_device(cuda):
    """test/test_PISCES.jl:32-92 through the device path (prescribed PAR, w = 0, Ω from the state update)."""
    grid = ob.RectilinearGrid(size=(1,), z=(-10, 0), topology=("Flat", "Flat", "Bounded"), device=cuda)
    PAR = {n: ob.CenterField(grid, n, 100.0) for n in ("PAR₁", "PAR₂", "PAR₃")}
    PAR["PAR"] = ob.CenterField(grid, "PAR", 300.0)
    bgc = ob.PISCES(grid, sinking_speeds={"POC": 0.0, "GOC": 0.0},
                    light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR),
                    mixed_layer_depth=ob.Field2D(grid, fill=-10.0), euphotic_depth=ob.Field2D(grid, fill=-10.0),
                    mean_mixed_layer_light=ob.Field2D(grid, fill=300.0),
                    iron=pisces.SimpleIron(excess_scavenging_enhancement=0.0),
                    nitrogen=pisces.NitrateAmmonia(maximum_fixation_rate=0.0))
    u = bgc.underlying_biogeochemistry
    model = ob.BiogeochemicalModel(grid, bgc)
    aux = bgc.biogeochemical_auxiliary_fields()
    # zero state → zero tendencies
    u.compute_tendencies(grid, model.tracers, aux, model.Gn, accumulate=False, time=0.0)
    assert all(bool((model.Gn[n].interior == 0).all()) for n in pisces.TRACERS[:24])
    model.set(**{n: v for n, v in pisces.PISCES_INITIAL_VALUES.items()})
    u.compute_tendencies(grid, model.tracers, aux, model.Gn, accumulate=False, time=1.0)
    G = {n: model.Gn[n].interior.item() for n in pisces.TRACERS}
    cons = u.conserved_tracers(ntuple=True)
    # the fused kernel contracts a·b + c into FMAs and divides with ≤ 1.5 ulp error, so budgets close to rounding
    # of the LARGEST term rather than to the reference's absolute 1e-20 (which the un-contracted oracle meets)
    for key in ("carbon", "silicon"):
        terms = [G[n] for n in cons[key]]
        assert abs(sum(terms)) <= 2e-15 * sum(abs(x) for x in terms)
    for key in ("iron", "phosphate", "nitrogen"):
        terms = [G[n] * f for n, f in zip(cons[key]["tracers"], cons[key]["scalefactors"])]
        assert abs(sum(terms)) <= 2e-15 * sum(abs(x) for x in terms)


def test_reference_conservation_test_as_a_box_model(cuda):
    """test/test_PISCES.jl:32-92 as written there: `BoxModelGrid(; z = -5)`, ConstantFields for zₘₓₗ, zₑᵤ, κ̄ and
    PAR̄ₘₓₗ, one `time_step!` from the zero state (everything but T, S stays 0), then the initial values, another
    `time_step!`, and the element budgets of `model.timestepper.Gⁿ`."""
    grid = ob.BoxModelGrid(1, device=cuda, z=-5)
    PAR = {n: ob.CenterField(grid, n, 100.0) for n in ("PAR₁", "PAR₂", "PAR₃")}
    PAR["PAR"] = ob.CenterField(grid, "PAR", 300.0)
    bgc = ob.PISCES(grid, sinking_speeds={"POC": 0.0, "GOC": 0.0},
                    light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR),
                    mixed_layer_depth=ob.ConstantField(grid, -10.0), euphotic_depth=ob.ConstantField(grid, -10.0),
                    mean_mixed_layer_vertical_diffusivity=ob.ConstantField(grid, 1.0),
                    mean_mixed_layer_light=ob.ConstantField(grid, 300.0),
                    iron=pisces.SimpleIron(excess_scavenging_enhancement=0.0),
                    nitrogen=pisces.NitrateAmmonia(maximum_fixation_rate=0.0))
    model = ob.BoxModel(biogeochemistry=bgc, grid=grid)
    model.time_step(1.0)
    assert all(f.interior.item() == 0.0 for n, f in model.fields.items() if n not in ("T", "S"))
    model.set(**pisces.PISCES_INITIAL_VALUES)
    before = {n: f.interior.item() for n, f in model.fields.items()}
    model.time_step(1.0)
    assert any(model.fields[n].interior.item() != before[n] for n in ("P", "NO₃", "DIC"))
    u = bgc.underlying_biogeochemistry
    assert u.euphotic_depth.data.item() == -10.0 and u.mean_mixed_layer_light.data.item() == 300.0  # left alone
    G = {n: model.Gn[n].interior.item() for n in pisces.TRACERS}
    cons = u.conserved_tracers(ntuple=True)
    for key in ("carbon", "silicon"):
        terms = [G[n] for n in cons[key]]
        assert abs(sum(terms)) <= 2e-15 * sum(abs(x) for x in terms), key
    for key in ("iron", "phosphate", "nitrogen"):
        terms = [G[n] * f for n, f in zip(cons[key]["tracers"], cons[key]["scalefactors"])]
        assert abs(sum(terms)) <= 2e-15 * sum(abs(x) for x in terms), key


def test_state_update_matches_oracle(cuda, oracle):
    """update_biogeochemical_state!: 3-band PAR from PChl + DChl, zₑᵤ, PAR̄ₘₓₗ, Ω (update_state.jl:1-17)."""
    grid, bgc, model = build(cuda, (40, 6, 30), (1e4, 1e3, 500.0))
    u = bgc.underlying_biogeochemistry
    host = fill(model, bgc)
    model.update_state()
    og = oracle.Grid.like(grid)
    la = bgc.light_attenuation
    bands, total = oracle.par_multiband(og, la.c_params(), host["PChl"], host["DChl"], 1.0, 120.0)
    rel = lambda a, b: float(np.max(np.abs(og.interior(a) - og.interior(b)) / np.abs(og.interior(b))))  # noqa: E731
    for n, name in enumerate(la.field_names):
        assert rel(la.fields[name].data.cpu().numpy(), bands[n]) <= RTOL_TENDENCY
    PARd = la.total.data.cpu().numpy()
    assert rel(PARd, total) <= RTOL_TENDENCY
    zeu = oracle.euphotic_depth(og, PARd)
    assert rel(u.euphotic_depth.data.cpu().numpy(), zeu) <= RTOL_TENDENCY
    mean = oracle.mixed_layer_mean(og, u.mixed_layer_depth.data.cpu().numpy(), PARd)
    assert rel(u.mean_mixed_layer_light.data.cpu().numpy(), mean) <= RTOL_TENDENCY
    Om = oracle.calcite_saturation(og, host["T"], host["S"], host["DIC"], host["Alk"], host["Si"])
    assert rel(u.calcite_saturation.data.cpu().numpy(), Om) <= RTOL_CARBON


def test_full_stage_runs_with_negative_scaling(cuda):
    grid = ob.RectilinearGrid(size=(16, 4, 10), extent=(1e3, 1e2, 100.0), device=cuda)
    bgc = ob.PISCES(grid, scale_negatives=True, surface_photosynthetically_active_radiation=80.0)
    model = ob.BiogeochemicalModel(grid, bgc)
    fill(model, bgc)
    # (a negative P would be zeroed and then θChl = PChl / (12·0 + eps) = Inf ⇒ 0·Inf = NaN in ∂ₜPChl — in the
    #  reference too: PChl is in no conserved group, PISCES/coupling_utils.jl:10 "TODO: deal with remaining")
    model.tracers["NO₃"].interior[0, 0, :4] = -0.1
    model.time_step(60.0)
    assert all(bool(torch.isfinite(f.interior).all()) for f in model.tracers.values())
    assert bool((model.tracers["NO₃"].interior >= 0).all())


@pytest.mark.parametrize("warm", [False, True])
def test_fused_scaling_and_calcite_saturation_equals_separate_launches(cuda, oracle, warm):
    """obm_scale_negative_tracers_calcite_saturation ≡ obm_scale_negative_tracers → obm_calcite_saturation: tracers bit
    for bit (incl. zeroed / NaN-filled cells), Ω from the RESCALED DIC, Alk, Si to rounding, and Ω ≡ oracle."""
    out = []
    for fuse in (True, False):
        grid = ob.RectilinearGrid(size=(37, 5, 21), extent=(1e4, 1e3, 2000.0), device=cuda)
        bgc = ob.PISCES(grid, scale_negatives=True, surface_photosynthetically_active_radiation=90.0)
        bgc.fuse_state_update = fuse
        bgc.underlying_biogeochemistry.warm_start_carbonate_solve = warm
        model = ob.BiogeochemicalModel(grid, bgc)
        fill(model, bgc)
        t = model.tracers
        t["DOC"].interior[3, 1, :7] = -5.0        # carbon group: DIC is rescaled (by ≈ 1e-3) in these cells
        t["Si"].interior[5:9, 2, 4] = -0.3        # silicon group: Si → 0, DSi / PSi rescaled
        t["NO₃"].interior[0, 0, 10:] = -0.1
        t["Fe"].interior[11, 3, 2] = float("nan")  # iron group only
        t["PSi"].interior[20, 4, 20] = -1e9       # group total < 0 ⇒ invalid_fill_value (NaN) on DSi, Si
        for _ in range(2 if warm else 1):          # second pass: the stored [H⁺] is used
            model.update_state()
        torch.cuda.synchronize()
        out.append((model, bgc))
    (mf, bf), (ms, bs) = out
    for n in pisces.TRACERS:
        a, b = mf.tracers[n].data, ms.tracers[n].data
        assert bool(((a == b) | (a.isnan() & b.isnan())).all()), n
    of, os_ = bf.underlying_biogeochemistry.calcite_saturation.interior, bs.underlying_biogeochemistry.calcite_saturation.interior
    both_nan = of.isnan() & os_.isnan()
    # the NaN-filled Si cell; a NaN Fe reaches Ω only on a second pass (iron group → NaN Z, M → carbon group → NaN DIC)
    assert int(both_nan.sum()) == (2 if warm else 1)
    assert bool((((of - os_).abs() <= 1e-13 * os_.abs()) | both_nan).all())
    og = oracle.Grid.like(mf.grid)
    h = {n: np.ascontiguousarray(mf.tracers[n].data.cpu().numpy()) for n in ("T", "S", "DIC", "Alk", "Si")}
    Om = og.interior(oracle.calcite_saturation(og, h["T"], h["S"], h["DIC"], h["Alk"], h["Si"]))
    got = of.cpu().numpy()
    ok = np.isfinite(Om) & np.isfinite(got)  # the NaN-filled cell is excluded (counted above)
    assert float(np.max(np.abs(got[ok] - Om[ok]) / np.abs(Om[ok]))) <= RTOL_CARBON
    for n in ("PAR", "PAR₁"):
        assert torch.equal(bf.biogeochemical_auxiliary_fields()[n].data, bs.biogeochemical_auxiliary_fields()[n].data)
    # obm_par_multiband_column_state ≡ obm_par_multiband → obm_euphotic_depth → obm_mixed_layer_mean
    uf, us = bf.underlying_biogeochemistry, bs.underlying_biogeochemistry
    zf_, zs_ = uf.euphotic_depth.interior, us.euphotic_depth.interior
    assert bool(((zf_ - zs_).abs() <= 1e-13 * zs_.abs()).all()) and bool((zs_ < -1.0).all())
    mf_, ms_ = uf.mean_mixed_layer_light.interior, us.mean_mixed_layer_light.interior
    assert bool(((mf_ - ms_).abs() <= 1e-13 * ms_.abs()).all()) and bool((ms_ > 0).all())


@pytest.mark.parametrize("size,extent", [((40, 6, 30), (1e4, 1e3, 500.0)), ((33, 3, 70), (1e3, 1e2, 4000.0)), ((5, 2, 7), (10.0, 4.0, 30.0))])
def test_par_scan_with_column_state_matches_oracle(cuda, oracle, size, extent):
    """The fused PAR launch: zₑᵤ and PAR̄ₘₓₗ against the oracle's top-down column loops, on columns where the euphotic
    depth falls in the first z-tile, in a later tile (70 levels, 4 km) and nowhere (7 shallow levels ⇒ znode(k = 0))."""
    grid, bgc, model = build(cuda, size, extent)
    assert bgc.fuse_state_update
    u = bgc.underlying_biogeochemistry
    fill(model, bgc)
    synthetic.fill_torch(u.mixed_layer_depth, "zₘₓₗ", -0.9 * extent[2], -0.02 * extent[2])
    model.update_state()
    og = oracle.Grid.like(grid)
    PARd = bgc.light_attenuation.total.data.cpu().numpy()
    zeu = oracle.euphotic_depth(og, PARd)
    got = u.euphotic_depth.data.cpu().numpy()
    assert float(np.max(np.abs(og.interior(got) - og.interior(zeu)) / np.abs(og.interior(zeu)))) <= RTOL_TENDENCY
    if size[2] == 7:
        assert np.all(og.interior(got) == grid.zc_host[grid.Hz - 1])  # never dark enough: znode(i, j, 0, grid, …)
    mean = oracle.mixed_layer_mean(og, u.mixed_layer_depth.data.cpu().numpy(), PARd)
    gm = u.mean_mixed_layer_light.data.cpu().numpy()
    assert float(np.max(np.abs(og.interior(gm) - og.interior(mean)) / np.abs(og.interior(mean)))) <= RTOL_TENDENCY


def test_per_tracer_call_form(cuda, oracle):
    """`bgc(i, j, k, grid, Val(name), clock, fields, auxiliary_fields)` (PISCES.jl:120-123): one tracer's tendency at a
    state is the fused kernel evaluated on boxes — against the oracle's point evaluation of the same state."""
    grid, bgc, model = build(cuda, (4, 2, 3), (10.0, 10.0, 30.0))
    u = bgc.underlying_biogeochemistry
    state = dict(pisces.PISCES_INITIAL_VALUES)
    state.update({"T": 14.0, "S": 35.0})
    aux = {"PAR₁": 31.0, "PAR₂": 22.0, "PAR₃": 9.0, "Ω": 0.8, "zₘₓₗ": -40.0, "zₑᵤ": -70.0, "κ": 1e-3, "mixed_layer_PAR": 25.0,
           "wPOC": -2 / 86400, "wGOC": -40 / 86400}
    t, z = 0.4 * 365 * 86400.0, -55.0
    want, S = oracle.pisces_point_terms(u.c_params(t), [state.get(n, 0.0) for n in pisces.TRACERS], aux["PAR₁"], aux["PAR₂"],
                                        aux["PAR₃"], 62.0, aux["Ω"], aux["wPOC"], aux["wGOC"], aux["zₘₓₗ"], aux["zₑᵤ"], aux["κ"],
                                        aux["mixed_layer_PAR"], z)
    for q, n in enumerate(pisces.TRACERS[:24]):
        got = u(n, z=z, time=t, device=cuda, **state, **aux)
        assert isinstance(got, float) and abs(got - want[q]) <= RTOL_TENDENCY * max(abs(want[q]), S[q]), (n, got, want[q])
    assert u("T", z=z, time=t, device=cuda, **state, **aux) == 0.0  # zero(grid), PISCES.jl:120
    Fe = torch.linspace(0.05, 1.5, 9, dtype=torch.float64)
    gP = u("PFe", z=z, time=t, device=cuda, **{**state, "Fe": Fe}, **aux)
    assert gP.shape == (9,) and bool((gP[1:] > gP[:-1]).all())  # iron uptake grows with dissolved iron


def test_model_latitude_rows_match_oracle_row_by_row(cuda, oracle):
    """`latitude = ModelLatitude()` on a LatitudeLongitudeGrid (PISCES/common.jl:27-28, PISCES.jl:360-367): every row of
    columns has its own latitude, hence its own day lengths — both of the reference's argument orders — and, south of
    the equator, the enhanced silicate limitation.  `obm_pisces_tendencies_rows` against the oracle called row by row
    with that row's scalar parameter block; and a table of identical rows ≡ the prescribed-latitude launch bit for bit."""
    from helpers import tendency_parity
    t = 0.61 * 365 * 86400.0
    grid = ob.LatitudeLongitudeGrid(size=(40, 6, 9), longitude=(0.0, 10.0), latitude=(-50.0, 58.0), z=(-300.0, 0.0), device=cuda)
    with pytest.warns(UserWarning, match="prescribed value is ignored"):
        ob.PISCES(grid, latitude=ob.PrescribedLatitude(10.0))
    bgc = ob.PISCES(grid, latitude=ob.ModelLatitude(), surface_photosynthetically_active_radiation=120.0)
    model = ob.BiogeochemicalModel(grid, bgc)
    u = bgc.underlying_biogeochemistry
    host = fill(model, bgc)
    model.clock.time = t
    model.update_state()
    og = oracle.Grid.like(grid)
    aux = host_aux(og, grid, bgc)
    G = {n: ob.CenterField(grid, fill=3.0) for n in pisces.TRACERS}
    u.compute_tendencies(grid, model.tracers, bgc.biogeochemical_auxiliary_fields(), G, accumulate=False, time=t)
    tr = [host[n] for n in pisces.TRACERS]
    Go = [np.zeros(og.parent_shape) if n < 24 else None for n in range(26)]
    So = [np.zeros(og.parent_shape) if n < 24 else None for n in range(26)]
    lats = [float(v) for v in grid.latitude_centers]
    assert min(lats) < 0 < max(lats)
    day_lengths = set()
    for j, lat in enumerate(lats):
        u.latitude = ob.PrescribedLatitude(lat)
        p = u.c_params(t)
        day_lengths.add((p.day_length_growth, p.day_length_chlorophyll))
        oracle.pisces_tendencies(og, p, tr, aux, G=Go, rows=(j, j + 1))
        oracle.pisces_tendency_scales(og, p, tr, aux, S=So, rows=(j, j + 1))
    assert len(day_lengths) == len(lats)  # the rows really differ
    worst = 0.0
    for n, name in enumerate(pisces.TRACERS[:24]):
        err, rel_max, rel_p = tendency_parity(og.interior(G[name].data.cpu().numpy()), og.interior(Go[n]), og.interior(So[n]))
        assert err <= RTOL_TENDENCY and rel_p <= 1e-12, (name, err, rel_max, rel_p)
        worst = max(worst, err)
    print(f"[parity] PISCES ModelLatitude, {len(lats)} rows: scale-aware max {worst:.2e}")
    # identical rows ≡ the scalar launch, bit for bit (same cell code, same derived values)
    u.latitude = ob.PrescribedLatitude(lats[1])
    G1 = {n: ob.CenterField(grid, fill=3.0) for n in pisces.TRACERS}
    u.compute_tendencies(grid, model.tracers, bgc.biogeochemical_auxiliary_fields(), G1, accumulate=False, time=t)
    u.latitude = ob.ModelLatitude()
    u.row_table = lambda g, time, _r=u.row_table: _r(g, time)[:, 1:2].expand(3, g.Ny).contiguous()
    G2 = {n: ob.CenterField(grid, fill=3.0) for n in pisces.TRACERS}
    u.compute_tendencies(grid, model.tracers, bgc.biogeochemical_auxiliary_fields(), G2, accumulate=False, time=t)
    for n in pisces.TRACERS[:24]:
        assert torch.equal(G1[n].data, G2[n].data), n
