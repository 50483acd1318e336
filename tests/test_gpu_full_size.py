"""GPU, BASELINE.json's full sizes: the whole stage through the hooks at LOBSTER 512×512×64 (configs[2]) and PISCES
1024×1024×128 (configs[3], 66 GB), checked through size-independent properties — element budgets of the tendencies
close in every cell, halos are never written, a second run reproduces the first bit for bit — and against the oracle
on a few thousand randomly drawn cells / columns (every kernel is pointwise or column-local, so a drawn cell or
column can be re-evaluated on the host from its own inputs)."""
import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import pisces, synthetic
from helpers import RTOL_TENDENCY, assert_tendency_parity

pytestmark = pytest.mark.gpu


def budget_residual(G, tracers, factors):
    """max over cells of |Σ f·G| / Σ |f·G| (0/0 → 0), computed on the device field by field."""
    tot = torch.zeros_like(G[tracers[0]].interior)
    mag = torch.zeros_like(tot)
    for n, f in zip(tracers, factors):
        g = G[n].interior
        tot.add_(g, alpha=float(f))
        mag.add_(g.abs(), alpha=abs(float(f)))
    r = tot.abs() / mag.clamp_min(1e-300)
    return r.max().item()


def draw_cells(grid, n, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, grid.Nz, n), rng.integers(0, grid.Ny, n), rng.integers(0, grid.Nx, n))


def at(field, k, j, i):
    v = field.interior
    kk = torch.as_tensor(k, device=v.device) if v.shape[0] > 1 else torch.zeros(len(i), dtype=torch.long, device=v.device)
    return v[kk, torch.as_tensor(j, device=v.device), torch.as_tensor(i, device=v.device)].cpu().numpy()


def eady_z_faces(Nz=64, Lz=140.0, refinement=1.8, stretching=3.0):
    """paper/figures/eady.jl:11-27 — the stretched vertical grid of the reference's Eady case (k = 1 … Nz + 1)."""
    def z(k):
        h = (k - 1) / Nz
        return Lz * ((1 + (h - 1) / refinement) * (1 - np.exp(-stretching * h)) / (1 - np.exp(-stretching)) - 1)
    return z


def test_lobster_c3_full_size(cuda, oracle):
    """BASELINE configs[2] as named: LOBSTER + carbonates + O₂ on the 512×512×64 Eady grid (stretched z of
    paper/figures/eady.jl) WITH the SimpleMultiG sediment bottom boundary under sinking sPOM / bPOM."""
    grid = ob.RectilinearGrid(size=(512, 512, 64), x=(0.0, 1000.0), y=(0.0, 1000.0), z=eady_z_faces(), device=cuda)
    assert abs(grid.zf[0] + 140.0) < 1e-9 and grid.dz[0] > 1.9 * grid.dz[-1]  # stretched: 3.06 m at the bottom, 1.56 m at the surface
    sed = ob.SimpleMultiGSediment(grid)
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen(), sediment=sed, scale_negatives=True,
                     surface_photosynthetically_active_radiation=100.0)
    model = ob.BiogeochemicalModel(grid, bgc, sinking_advection="UpwindBiased1")
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *synthetic.lobster_range(n))
    for n, f in sed.fields.items():
        synthetic.fill_torch(f, "sed" + n, 1e-2, 10.0, True)
    model.clock.last_stage_dt = 60.0
    sed.last_dt = 60.0
    for n, f in sed.Gn.items():   # the tendencies the sediment holds from "the previous stage"
        synthetic.fill_torch(f, "sedG" + n, -1e-6, 1e-6)
    pools_before = {n: f.data.clone() for n, f in sed.fields.items()}
    sedG_before = {n: f.data.clone() for n, f in sed.Gn.items()}
    model.update_state()          # negative scaling, two-band PAR, sediment: tracked fields + AB2 step + new pool tendencies
    model.compute_tendencies()    # fused tendencies, sediment ↔ bottom-cell fluxes, upwind sinking of sPOM / bPOM
    torch.cuda.synchronize()
    u = bgc.underlying_biogeochemistry
    groups = u.conserved_tracers(labeled=True)
    # --- nitrogen: water column + sediment, column by column.  Σ_k Σ_N Gⁿ·Δz_k (BGC sources sum to zero, sinking
    # telescopes to the open bottom face, the sediment returns NO₃ / NH₄ to the bottom cell) + Σ_pools Gⁿ_sed = 0 — the
    # property the reference's (commented-out) sediment test intends, test/test_sediments.jl:37-80
    dz = torch.from_numpy(grid.dz).to(cuda).reshape(-1, 1, 1)
    colN = torch.zeros(grid.Ny, grid.Nx, dtype=torch.float64, device=cuda)
    magN = torch.zeros_like(colN)
    for n in groups["nitrogen"]:
        g = model.Gn[n].interior * dz
        colN += g.sum(0)
        magN += g.abs().sum(0)
    pools = torch.zeros_like(colN)
    for n in ("Ns", "Nf", "Nr"):
        pools += sed.Gn[n].interior[0]
        magN += sed.Gn[n].interior[0].abs()
    # what crosses the bottom face of the water column is what the sediment receives, and what the pools lose returns to
    # the bottom cell as NO₃ + NH₄ (simple_multi_G.jl:165-226): every column's budget closes to rounding
    flux_in = sum(sed.tracked_fields[n].interior[0] for n in sed.biogeochemistry.sinking_fluxes())
    assert bool((flux_in > 0).all())
    resid = (colN + pools).abs() / magN
    assert float(resid.max()) <= RTOL_TENDENCY, float(resid.max())
    # the sediment stepped its pools with the tendencies it held, and halos / other planes of Gⁿ are untouched
    for n in sed.fields:
        want = pools_before[n] + 60.0 * (1.5 + 0.1) * sedG_before[n]  # AB2, χ = 0.1, G⁻ = 0 (timesteppers.jl:29-73)
        assert torch.allclose(grid.interior(sed.fields[n].data), grid.interior(want), rtol=1e-13, atol=0)
        assert torch.equal(grid.interior(sed.Gm[n].data), grid.interior(sedG_before[n]))  # cached: G⁻ ← Gⁿ
    g = model.Gn["NO₃"].data
    assert g.sum().item() == model.Gn["NO₃"].interior.sum().item()
    first = {n: f.data.clone() for n, f in model.Gn.items()}
    # --- 3000 drawn cells of the fused tendencies against the oracle (BGC part: recompute without sinking / sediment)
    plain = ob.BiogeochemicalModel(grid, ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen(),
                                                    surface_photosynthetically_active_radiation=100.0))
    plain.tracers = model.tracers
    plain.biogeochemistry.light_attenuation = bgc.light_attenuation
    for gfield in plain.Gn.values():
        gfield.data.zero_()
    plain.biogeochemistry.update_tendencies(plain)
    k, j, i = draw_cells(grid, 3000, 1)
    names = list(u.required_biogeochemical_tracers())
    og = oracle.Grid(3000, 1, 1, 0, 0, 0, np.array([-0.5]), np.array([-1.0, 0.0]))
    vals = [np.ascontiguousarray(at(model.tracers[n], k, j, i).reshape(og.parent_shape)) for n in names]
    PAR = np.ascontiguousarray(at(bgc.biogeochemical_auxiliary_fields()["PAR"], k, j, i).reshape(og.parent_shape))
    want = dict(zip(names, oracle.npd_tendencies(og, u.c_params(), vals, PAR)))
    S = dict(zip(names, oracle.npd_tendency_scales(og, u.c_params(), vals, PAR)))  # Σ|terms| per tendency
    got = {n: at(plain.Gn[n], k, j, i).reshape(og.parent_shape) for n in names}
    assert_tendency_parity("lobster_c3_full_size_3000_cells", names, got, want, S)
    del plain
    # --- 32 drawn columns of the sediment hooks against the oracle: pools after the AB2 step, new pool tendencies,
    # tracked fields, and the fluxes added to the bottom cells of NO₃, NH₄, O₂, DIC
    rng = np.random.default_rng(5)
    cj, ci = rng.integers(0, grid.Ny, 32), rng.integers(0, grid.Nx, 32)
    ogc = oracle.Grid(32, 1, grid.Nz, 0, 0, grid.Hz, grid.zc_host, grid.zf_host)
    tj, ti = torch.as_tensor(cj + grid.Hy, device=cuda), torch.as_tensor(ci + grid.Hx, device=cuda)

    def columns(data):  # parent (Nz + 2Hz | 1, Ny + 2Hy, Nx + 2Hx) → host parent array of the 32-column grid
        out = np.zeros((data.shape[0], 1, 32))
        out[:, 0, :] = data[:, tj, ti].cpu().numpy()
        return out

    b = sed.biogeochemistry
    h_tr = {n: columns(model.tracers[n].data) for n in ("NO₃", "NH₄", "O₂", "sPOM", "bPOM")}
    h_w = {n: columns(sed._w_field(bgc, n).data) for n in b.sinking_fluxes()}
    h_pools = [columns(pools_before[n]) for n in sed.fields]
    h_Gn = [columns(sedG_before[n]) for n in sed.fields]
    h_Gm = [np.zeros(ogc.plane_shape) for _ in sed.fields]
    h_tracked = [np.zeros(ogc.plane_shape) for _ in sed.tracked_fields]
    h_Gc = [np.zeros(ogc.parent_shape) for _ in b.coupled_tracers()]
    fo = oracle.sediment_fields(NO3=h_tr["NO₃"], NH4=h_tr["NH₄"], O2=h_tr["O₂"], sinking=[h_tr[n] for n in b.sinking_fluxes()],
                                sinking_w=[h_w[n] for n in b.sinking_fluxes()], pools=h_pools, Gn=h_Gn, Gm=h_Gm,
                                tracked=h_tracked, G_coupled=h_Gc)
    oracle.sediment_update_state(ogc, sed.c_params(), fo, 60.0, chi=0.1)
    oracle.sediment_update_tendencies(ogc, sed.c_params(), fo)
    close = lambda a, b_: np.testing.assert_allclose(a, b_, rtol=1e-12, atol=1e-300)  # noqa: E731
    for f, h in zip(sed.fields.values(), h_pools):
        close(columns(f.data), h)
    for f, h in zip(sed.Gn.values(), h_Gn):
        close(columns(f.data), h)
    for f, h in zip(sed.tracked_fields.values(), h_tracked):
        close(columns(f.data), h)
    # a second pass from the same inputs reproduces the tendencies bit for bit (the sediment pools are restored first)
    for n in sed.fields:
        sed.fields[n].data.copy_(pools_before[n]); sed.Gn[n].data.copy_(sedG_before[n]); sed.Gm[n].data.zero_()
    sed.iteration, sed.last_dt = 0, 60.0
    model.update_state()
    model.compute_tendencies()
    assert all(torch.equal(first[n], model.Gn[n].data) for n in first)
    # … and 24 drawn columns of the two-band PAR scan on the stretched grid
    rng = np.random.default_rng(2)
    cj, ci = rng.integers(0, grid.Ny, 24), rng.integers(0, grid.Nx, 24)
    ogc = oracle.Grid(24, 1, grid.Nz, 0, 0, grid.Hz, grid.zc_host, grid.zf_host)
    P = np.zeros(ogc.parent_shape)
    Pdev = model.tracers["P"].data[:, torch.as_tensor(cj + grid.Hy, device=cuda), torch.as_tensor(ci + grid.Hx, device=cuda)]
    P[:, 0, :] = Pdev.cpu().numpy()
    wantPAR = oracle.par_twoband(ogc, bgc.light_attenuation.c_params(), P, 100.0)
    gotPAR = bgc.biogeochemical_auxiliary_fields()["PAR"].interior[:, torch.as_tensor(cj, device=cuda),
                                                                  torch.as_tensor(ci, device=cuda)].cpu().numpy()
    wp = ogc.interior(wantPAR)[:, 0, :]
    assert np.max(np.abs(gotPAR - wp) / np.abs(wp)) <= 1e-12


def test_pisces_c4_full_size(cuda, oracle):
    free, _ = torch.cuda.mem_get_info(cuda)
    if free < 75e9:
        pytest.skip(f"needs ≈ 70 GB of device memory, {free / 1e9:.0f} GB free")
    grid = ob.RectilinearGrid(size=(1024, 1024, 128), extent=(10240.0, 10240.0, 400.0), device=cuda)
    bgc = ob.PISCES(grid, scale_negatives=True, surface_photosynthetically_active_radiation=100.0)
    model = ob.BiogeochemicalModel(grid, bgc)
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *pisces.synthetic_range(n))
    pisces.fill_synthetic_auxiliary(bgc, model)
    model.update_state()
    model.compute_tendencies()
    torch.cuda.synchronize()
    u = bgc.underlying_biogeochemistry
    groups = u.conserved_tracers(ntuple=True)
    # carbon, silicon and phosphorus are closed by the source terms alone (nitrogen has fixation, iron has scavenging)
    assert budget_residual(model.Gn, groups["carbon"], [1] * 9) <= RTOL_TENDENCY
    assert budget_residual(model.Gn, groups["silicon"], [1] * 3) <= RTOL_TENDENCY
    assert budget_residual(model.Gn, groups["phosphate"]["tracers"], groups["phosphate"]["scalefactors"]) <= RTOL_TENDENCY
    for n in ("T", "S"):
        assert not model.Gn[n].data.any()  # zero(grid), PISCES.jl:120
    g = model.Gn["DIC"].data
    assert g.sum().item() == model.Gn["DIC"].interior.sum().item()  # halos untouched
    assert all(torch.isfinite(model.Gn[n].interior).all() for n in ("P", "Fe", "O₂", "Alk"))
    # 4000 drawn cells against the oracle's point evaluation, inputs gathered from the device state
    k, j, i = draw_cells(grid, 4000, 3)
    aux = bgc.biogeochemical_auxiliary_fields()
    T = {n: at(model.tracers[n], k, j, i) for n in pisces.TRACERS}
    A = {n: at(aux[n], k, j, i) for n in ("PAR₁", "PAR₂", "PAR₃", "PAR", "Ω", "zₘₓₗ", "zₑᵤ", "κ", "mixed_layer_PAR")}
    kt, jt, it = (torch.as_tensor(x, device=cuda) for x in (k, j, i))

    def wmean(f):  # ℑzᵃᵃᶜ of a z-face field: (w[k] + w[k+1]) / 2
        d = f.data
        return ((d[kt + grid.Hz, jt + grid.Hy, it + grid.Hx] + d[kt + grid.Hz + 1, jt + grid.Hy, it + grid.Hx]) / 2).cpu().numpy()

    wP, wG = wmean(aux["wPOC"]), wmean(aux["wGOC"])
    G = {n: at(model.Gn[n], k, j, i) for n in pisces.TRACERS[:24]}
    params = u.c_params(model.clock.time)
    names24 = pisces.TRACERS[:24]
    want = {n: np.zeros(len(k)) for n in names24}
    S = {n: np.zeros(len(k)) for n in names24}
    for c in range(len(k)):
        w, sc = oracle.pisces_point_terms(params, [T[n][c] for n in pisces.TRACERS], A["PAR₁"][c], A["PAR₂"][c], A["PAR₃"][c],
                                          A["PAR"][c], A["Ω"][c], wP[c], wG[c], A["zₘₓₗ"][c], A["zₑᵤ"][c], A["κ"][c],
                                          A["mixed_layer_PAR"][c], grid.zc[k[c]])
        for q, n in enumerate(names24):
            want[n][c], S[n][c] = w[q], sc[q]
    # the stated metric: |Δ| ≤ 1e-12·max(|want|, S) with S = Σ|additive terms| of THAT tendency; pure relative error too
    assert_tendency_parity("pisces_c4_full_size_4000_cells", names24, G, want, S)
    # Ω of the drawn cells against the reference's own damped Newton (oracle), P = |z|·g·1026/1e5 bar
    om = np.array([oracle.carbon_chemistry(T["DIC"][c], T["T"][c], T["S"][c], T["Alk"][c],
                                           P=abs(grid.zc[k[c]]) * 9.80665 * 1026.0 / 100000.0, silicate=T["Si"][c],
                                           output=ob._lib.CC_OMEGA_CALCITE) for c in range(0, len(k), 8)])
    assert np.max(np.abs(A["Ω"][::8] - om) / np.abs(om)) <= 1e-10
    del model, bgc
    torch.cuda.empty_cache()
