"""GPU: parameter-sweep ensembles of the NPD family — `obm_npd_tendencies_ensemble` through the host
mirror.  Each horizontal column of the grid is a member with its own value of the varied parameters; the reference
gets the same thing by building and running one model per parameter vector (examples/data_assimilation.jl:26-52,
118-130), which is exactly how the checks below are assembled: member m of the ensemble launch ≡ the oracle (and the
single-parameter-set kernel) run with member m's parameter block."""
import math

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib, synthetic
from helpers import RTOL_TENDENCY, scale_aware_error, synthetic_state

pytestmark = pytest.mark.gpu

day = 86400.0
minutes = 60.0


def member_block(base, varied, m):
    """obm_npd_params of member m: a copy of `base` with the member's values."""
    p = _lib.obm_npd_params.from_buffer_copy(bytes(base))
    for name, v in varied.items():
        setattr(p, name, float(v[m]))
    return p


def sweep(rng, members, spec):
    return {name: centre * rng.uniform(lo, hi, members) for name, (centre, lo, hi) in spec.items()}


LOBSTER_SWEEP = {  # centre value, multiplicative range
    "phytoplankton_maximum_growth_rate": (1.21e-5, 0.5, 1.5),
    "nitrate_half_saturation": (0.7, 0.5, 2.0),
    "light_half_saturation": (33.0, 0.5, 2.0),
    "maximum_grazing_rate": (9.26e-6, 0.5, 1.5),
    "small_remineralisation_rate": (5.88e-7, 0.2, 3.0),
    "redfield_ratio": (6.56, 0.9, 1.1),
    "nitrification_rate": (5.8e-7, 0.0, 2.0),
    "respiration_oxygen_nitrogen_ratio": (10.75, 0.9, 1.1),
}


@pytest.mark.parametrize("accumulate", [False, True])
def test_every_member_matches_the_oracle_run_with_its_parameters(cuda, oracle, accumulate):
    grid = ob.RectilinearGrid(size=(13, 3, 5), extent=(13, 3, 50), device=cuda)  # 39 members × 5 levels
    members = grid.Nx * grid.Ny
    rng = np.random.default_rng(11)
    varied = sweep(rng, members, LOBSTER_SWEEP)
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen(),
                     parameter_ensemble=varied).underlying_biogeochemistry
    names = bgc.required_biogeochemical_tracers()
    dev, host, og = synthetic_state(grid, names, synthetic.lobster_range)
    pdev, phost, _ = synthetic_state(grid, ["PAR"], {"PAR": (0.0, 150.0, False)})
    g0 = 1e-7 if accumulate else 7.0
    G = {n: ob.CenterField(grid, "G" + n, fill=g0) for n in names}
    bgc.compute_tendencies(grid, dev, pdev, G, accumulate=accumulate)
    torch.cuda.synchronize()
    got = {n: og.interior(G[n].data.cpu().numpy()) for n in names}
    base = bgc.c_params()
    worst = 0.0
    for m in range(members):
        i, j = m % grid.Nx, m // grid.Nx
        Go = oracle.npd_tendencies(og, member_block(base, varied, m), [host[n] for n in names], phost["PAR"],
                                   G=[np.full(og.parent_shape, g0) for _ in names] if accumulate else None,
                                   accumulate=accumulate)
        want = {n: og.interior(g)[:, j, i] for n, g in zip(names, Go)}
        off = g0 if accumulate else 0.0
        S = np.maximum.reduce([np.abs(want[n] - off) for n in names]) + off
        for n in names:
            worst = max(worst, scale_aware_error(got[n][:, j, i], want[n], S))
    assert worst <= RTOL_TENDENCY, f"scale-aware error {worst:.3e}"
    # the members really differ: the same state under member 0's and member 1's parameters
    assert not np.array_equal(got["P"][:, 0, 0], got["P"][:, 0, 1])


def test_uniform_sweep_equals_the_plain_kernel(cuda):
    """Every member given the default values: the ensemble instantiation reproduces the plain kernel — to rounding, not
    bit for bit (the compiler is free to contract different multiply-adds in the two instantiations)."""
    grid = ob.RectilinearGrid(size=(37, 5, 9), extent=(37, 5, 90), device=cuda)
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen()).underlying_biogeochemistry
    names = bgc.required_biogeochemical_tracers()
    dev, _, _ = synthetic_state(grid, names, synthetic.lobster_range)
    pdev, _, _ = synthetic_state(grid, ["PAR"], {"PAR": (0.0, 150.0, False)})
    G1 = {n: ob.CenterField(grid) for n in names}
    G2 = {n: ob.CenterField(grid) for n in names}
    bgc.compute_tendencies(grid, dev, pdev, G1, accumulate=False)
    members = grid.Nx * grid.Ny
    bgc.set_parameter_ensemble(maximum_grazing_rate=np.full(members, bgc.plankton.maximum_grazing_rate),
                               nitrification_rate=np.full(members, bgc.nutrients.nitrification_rate))
    bgc.compute_tendencies(grid, dev, pdev, G2, accumulate=False)
    torch.cuda.synchronize()
    g1 = {n: G1[n].interior.cpu().numpy() for n in names}
    g2 = {n: G2[n].interior.cpu().numpy() for n in names}
    S = np.maximum.reduce([np.abs(g1[n]) for n in names])
    assert max(scale_aware_error(g2[n], g1[n], S) for n in names) <= 1e-14


def test_a_member_does_not_depend_on_its_neighbours(cuda):
    """Member m of a 39-member launch ≡ the same member alone on a one-column grid, bit for bit (same instantiation)."""
    grid = ob.RectilinearGrid(size=(13, 3, 5), extent=(13, 3, 50), device=cuda)
    members = grid.Nx * grid.Ny
    varied = sweep(np.random.default_rng(12), members, LOBSTER_SWEEP)
    bgc = ob.LOBSTER(grid, parameter_ensemble=varied).underlying_biogeochemistry
    names = bgc.required_biogeochemical_tracers()
    dev, _, _ = synthetic_state(grid, names, synthetic.lobster_range)
    pdev, _, _ = synthetic_state(grid, ["PAR"], {"PAR": (0.0, 150.0, False)})
    G = {n: ob.CenterField(grid) for n in names}
    bgc.compute_tendencies(grid, dev, pdev, G, accumulate=False)
    for m in (0, 17, 38):
        i, j = m % grid.Nx, m // grid.Nx
        col = ob.RectilinearGrid(size=(1, 1, 5), extent=(1, 1, 50), device=cuda)
        one = ob.LOBSTER(col, parameter_ensemble={k: v[m:m + 1] for k, v in varied.items()}).underlying_biogeochemistry
        f1 = {n: ob.CenterField(col, n) for n in names}
        for n in names:
            f1[n].interior[:, 0, 0] = dev[n].interior[:, j, i]
        P1 = ob.CenterField(col, "PAR")
        P1.interior[:, 0, 0] = pdev["PAR"].interior[:, j, i]
        G1 = {n: ob.CenterField(col) for n in names}
        one.compute_tendencies(col, f1, {"PAR": P1}, G1, accumulate=False)
        torch.cuda.synchronize()
        for n in names:
            assert torch.equal(G1[n].interior[:, 0, 0], G[n].interior[:, j, i]), (m, n)


def test_bad_parameter_index_and_count_are_refused(cuda):
    import ctypes as C
    grid = ob.BoxModelGrid(4, device=cuda)
    bgc = ob.NPZD(grid).underlying_biogeochemistry
    names = bgc.required_biogeochemical_tracers()
    f = {n: ob.CenterField(grid, n, 1.0) for n in names}
    G = {n: ob.CenterField(grid) for n in names}
    PAR = ob.CenterField(grid, "PAR", 10.0)
    lib = _lib.load()
    cg, p = grid.c_grid(), bgc.c_params()
    tptr = _lib.pointer_table([f[n].ptr for n in names])
    gptr = _lib.pointer_table([G[n].ptr for n in names])
    vals = torch.ones(4, dtype=torch.float64, device=cuda)
    rc = lib.obm_npd_tendencies_ensemble(C.byref(cg), C.byref(p), 1, (C.c_int32 * 1)(35), vals.data_ptr(), tptr, PAR.ptr,
                                         gptr, 0, None)
    assert rc == -3 and b"not a parameter index" in lib.obm_last_error()
    rc = lib.obm_npd_tendencies_ensemble(C.byref(cg), C.byref(p), 17, (C.c_int32 * 17)(), vals.data_ptr(), tptr, PAR.ptr,
                                         gptr, 0, None)
    assert rc != 0 and b"nvary" in lib.obm_last_error()
    rc = lib.obm_npd_tendencies_ensemble(C.byref(cg), C.byref(p), 1, (C.c_int32 * 1)(0), None, tptr, PAR.ptr, gptr, 0, None)
    assert rc != 0 and b"NULL" in lib.obm_last_error()


# ---- the calibration loop of examples/data_assimilation.jl, all members in one model --------------------------------

def PAR_func(t):  # examples/data_assimilation.jl:22-25
    year = 365 * day
    PAR0 = 60 * (1 - math.cos((t + 15 * day) * 2 * math.pi / year)) \
        * (1 / (1 + 0.2 * math.exp(-(((t % year) - 200 * day) / (50 * day)) ** 2))) + 2
    return PAR0 * math.exp(0.2 * -10)


def calibration_parameters(u):
    """run_box_simulation's PhytoZoo keywords from (α, μ₀, k_N, m_P) — data_assimilation.jl:39-43."""
    alpha, mu, kN, mP = u
    return {"phytoplankton_maximum_growth_rate": mu, "nitrate_half_saturation": kN, "light_half_saturation": mu / alpha,
            "phytoplankton_mortality_rate": 0.066 / day + mP / day,
            "phytoplankton_solid_waste_fraction": mP * day / (0.066 + mP * day)}


def calibration_model(cuda, n, **kw):
    grid = ob.BoxModelGrid(n, device=cuda)
    PAR = ob.CenterField(grid, "PAR")
    bgc = ob.NPZD(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR), **kw)
    model = ob.BoxModel(biogeochemistry=bgc, grid=grid, prescribed_tracers={"PAR": PAR_func, "T": lambda t: 12.0})
    model.set(N=10.0, P=0.1, Z=0.01)
    return model


def test_calibration_ensemble_equals_one_box_model_per_parameter_vector(cuda):
    """Eight members (N_ensemble, data_assimilation.jl:113) stepped together, eagerly and as a replayed CUDA graph,
    against separately built single-box models: bit for bit when the lone member takes the same (ensemble) kernel
    instantiation, to rounding when it is built the reference's way, from its own parameter vector."""
    rng = np.random.default_rng(41)
    n, steps = 8, 60
    u = np.stack([rng.normal(0.1953, 0.05, n).clip(0.05) / day, rng.normal(0.6989, 0.1, n).clip(0.1) / day,
                  rng.normal(2.3868, 0.5, n).clip(0.5), rng.normal(0.0101, 0.01, n).clip(1e-3) / day])
    per_member = [calibration_parameters(u[:, m]) for m in range(n)]
    ensemble = {k: np.array([pm[k] for pm in per_member]) for k in per_member[0]}

    eager = calibration_model(cuda, n, parameter_ensemble=ensemble)
    graph = calibration_model(cuda, n, parameter_ensemble=ensemble)
    for _ in range(steps):
        eager.time_step(20 * minutes)
    graph.run(20 * minutes, steps, graph=True)
    torch.cuda.synchronize()
    for name in eager.prognostic:
        assert torch.equal(eager.fields[name].data, graph.fields[name].data), name

    final_P = []
    for m in (0, 3, 7):
        pm = per_member[m]
        # the same member alone (same kernel instantiation): bit for bit
        alone = calibration_model(cuda, 1, parameter_ensemble={k: v[m:m + 1] for k, v in ensemble.items()})
        # and the way the reference does it — a model built from that parameter vector (plain kernel): to rounding
        one = calibration_model(cuda, 1, plankton=npzd_plankton(**pm))
        for _ in range(steps):
            alone.time_step(20 * minutes)
            one.time_step(20 * minutes)
        for name in one.prognostic:
            want = eager.fields[name].interior.reshape(-1)[m].item()
            assert alone.fields[name].interior.item() == want, (m, name)
            assert abs(one.fields[name].interior.item() - want) <= 1e-11 * max(abs(want), 1e-3), (m, name)
        final_P.append(one.fields["P"].interior.item())
    assert len(set(final_P)) == 3  # the parameters matter


def npzd_plankton(**override):
    """The PhytoZoo of `NPZD(grid)` (constructors.jl:177-227) with some keywords replaced."""
    base = ob.NPZD(ob.BoxModelGrid(1, device="cpu")).underlying_biogeochemistry.plankton
    import dataclasses
    return dataclasses.replace(base, **override)
