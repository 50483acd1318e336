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


def test_lobster_c3_full_size(cuda, oracle):
    grid = ob.RectilinearGrid(size=(512, 512, 64), extent=(5120.0, 5120.0, 140.0), device=cuda)
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen(), scale_negatives=True,
                     surface_photosynthetically_active_radiation=100.0)
    model = ob.BiogeochemicalModel(grid, bgc)
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *synthetic.lobster_range(n))
    model.update_state()
    model.compute_tendencies()
    torch.cuda.synchronize()
    u = bgc.underlying_biogeochemistry
    groups = u.conserved_tracers(labeled=True)
    assert budget_residual(model.Gn, groups["nitrogen"], [1] * len(groups["nitrogen"])) <= RTOL_TENDENCY
    assert budget_residual(model.Gn, groups["carbon"]["tracers"], groups["carbon"]["scalefactors"]) <= RTOL_TENDENCY
    # halos of Gⁿ untouched (still the zeros of compute_tendencies), second run bit-identical
    g = model.Gn["NO₃"].data
    assert g.sum().item() == model.Gn["NO₃"].interior.sum().item()
    first = {n: f.data.clone() for n, f in model.Gn.items()}
    model.update_state()
    model.compute_tendencies()
    assert all(torch.equal(first[n], model.Gn[n].data) for n in first)
    # 3000 drawn cells re-evaluated by the oracle from their own inputs (PAR taken from the device scan) …
    k, j, i = draw_cells(grid, 3000, 1)
    names = list(u.required_biogeochemical_tracers())
    og = oracle.Grid(3000, 1, 1, 0, 0, 0, np.array([-0.5]), np.array([-1.0, 0.0]))
    vals = [np.ascontiguousarray(at(model.tracers[n], k, j, i).reshape(og.parent_shape)) for n in names]
    PAR = np.ascontiguousarray(at(bgc.biogeochemical_auxiliary_fields()["PAR"], k, j, i).reshape(og.parent_shape))
    want = dict(zip(names, oracle.npd_tendencies(og, u.c_params(), vals, PAR)))
    S = dict(zip(names, oracle.npd_tendency_scales(og, u.c_params(), vals, PAR)))  # Σ|terms| per tendency
    got = {n: at(model.Gn[n], k, j, i).reshape(og.parent_shape) for n in names}
    assert_tendency_parity("lobster_c3_full_size_3000_cells", names, got, want, S)
    # … and 24 drawn columns of the two-band PAR scan
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
