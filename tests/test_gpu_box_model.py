"""GPU: the fused tracer update (obm_rk3_substep, bit-exact vs the oracle) and the device-resident BoxModel
ensemble driver — against a CPU loop assembled from the oracle in the reference's call order
(boxmodel.jl:92-110, timesteppers.jl:30-93), the behaviour the reference's own test checks
(test/test_boxmodel.jl:38-53: 10 steps of 20 minutes change every field), ensemble members ≡ single boxes,
and CUDA-graph replay ≡ eager stepping bit for bit."""
import math

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from helpers import RTOL_TENDENCY

pytestmark = pytest.mark.gpu

day = 86400.0
minutes = 60.0


def PAR_fn(t):  # test/test_boxmodel.jl:9
    return (60 * (1 - math.cos((t + 15 * day) * 2 * math.pi / (365 * day)))
            * (1 / (1 + 0.2 * math.exp(-(((t % (365 * day)) - 200 * day) / (50 * day)) ** 2))) + 2) * math.exp(-2)


DEFAULTS = {"NO₃": 10.0, "NH₄": 0.1, "P": 0.1, "Z": 0.01}


def simple_box_model(cuda, n=1, **kw):
    grid = ob.BoxModelGrid(n, device=cuda)
    PAR = ob.CenterField(grid, "PAR")
    bgc = ob.LOBSTER(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR), **kw)
    return ob.BoxModel(biogeochemistry=bgc, grid=grid, prescribed_tracers={"PAR": PAR_fn}), PAR


def host(f):
    return np.ascontiguousarray(f.data.cpu().numpy())


def test_substep_bit_exact(cuda, oracle):
    grid = ob.RectilinearGrid(size=(70, 5, 6), extent=(1, 1, 1), device=cuda)
    og = oracle.Grid.like(grid)
    rng = np.random.default_rng(0)
    nf = 7
    hU, hGn, hGm = ([rng.normal(size=og.parent_shape) for _ in range(nf)] for _ in range(3))
    mk = lambda hs: [ob.CenterField(grid).set(og.interior(h)) for h in hs]  # noqa: E731
    from oceanbiome_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    for gamma, zeta, cache in ((8 / 15, None, 1), (5 / 12, -17 / 60, 1), (1.0, None, 0)):
        dU, dGn, dGm = mk(hU), mk(hGn), mk(hGm)
        for d, h in zip(dU + dGn + dGm, hU + hGn + hGm):
            d.data.copy_(torch.from_numpy(h))  # halos too
        cg = grid.c_grid()
        tab = lambda fs: _lib.pointer_table([f.ptr for f in fs])  # noqa: E731
        rc = lib.obm_rk3_substep(C.byref(cg), nf, tab(dU), tab(dGn), tab(dGm), 1200.0, gamma, zeta or 0.0,
                                 int(zeta is not None), cache, None)
        assert rc == 0
        wU, wGm = [u.copy() for u in hU], [m.copy() for m in hGm]
        oracle.rk3_substep(og, wU, hGn, wGm, 1200.0, gamma, zeta, cache_previous=bool(cache))
        for f in range(nf):
            assert np.array_equal(host(dU[f]), wU[f]) and np.array_equal(host(dGm[f]), wGm[f])
    # argument errors
    assert lib.obm_rk3_substep(C.byref(cg), 3, None, None, None, 1.0, 1.0, 0.0, 0, 0, None) == -1
    assert lib.obm_rk3_substep(C.byref(cg), -1, None, None, None, 1.0, 1.0, 0.0, 0, 0, None) == -2
    assert lib.obm_rk3_substep(C.byref(cg), 0, None, None, None, 1.0, 1.0, 0.0, 0, 0, None) == 0


def test_reference_behaviour_ten_steps(cuda):
    # test/test_boxmodel.jl:38-53
    model, _ = simple_box_model(cuda)
    model.set(**DEFAULTS)
    for _ in range(10):
        model.time_step(20 * minutes)
    for name, f in model.fields.items():
        assert f.interior.item() != DEFAULTS.get(name, 0.0), name
        assert math.isfinite(f.interior.item())
    assert model.clock.iteration == 10 and abs(model.clock.time - 10 * 20 * minutes) < 1e-6
    assert "box model" in model.summary()
    with pytest.raises(ValueError):
        model.set(nope=1.0)


def test_matches_cpu_loop_built_from_the_oracle(cuda, oracle):
    model, _ = simple_box_model(cuda, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen())
    ic = dict(DEFAULTS, sPOM=0.2, bPOM=0.1, DOM=0.3, DIC=2200.0, Alk=2400.0, **{"O₂": 240.0})
    model.set(**ic)
    names = list(model.biogeochemistry.required_biogeochemical_tracers())
    og = oracle.Grid.like(model.grid)
    params = model.biogeochemistry.underlying_biogeochemistry.c_params()
    U = [np.full(og.parent_shape, ic.get(n, 0.0)) for n in names]
    Gn = [np.zeros(og.parent_shape) for _ in names]
    Gm = [np.zeros(og.parent_shape) for _ in names]
    dt, t = 20 * minutes, 0.0

    def tendencies(tt):
        return oracle.npd_tendencies(og, params, U, np.full(og.parent_shape, PAR_fn(tt)))

    Gn = tendencies(t)  # update_state! at iteration 0
    for _ in range(25):
        for gamma, zeta in ob.BoxModel.RK3:
            oracle.rk3_substep(og, U, Gn, Gm, dt, gamma, zeta, cache_previous=True)
            t += dt * (gamma + (zeta or 0.0))
            Gn = tendencies(t)
        model.time_step(dt)
    for n, u in zip(names, U):
        got, want = model.fields[n].interior.item(), og.interior(u).item()
        assert abs(got - want) <= 50 * RTOL_TENDENCY * max(abs(want), 1e-3), (n, got, want)


def test_ensemble_members_are_independent_boxes(cuda):
    n = 257
    rng = np.random.default_rng(3)
    ics = {k: v * rng.uniform(0.5, 1.5, n) for k, v in DEFAULTS.items()}
    ens, _ = simple_box_model(cuda, n)
    ens.set(**{k: torch.from_numpy(v) for k, v in ics.items()})
    for _ in range(5):
        ens.time_step(20 * minutes)
    for m in (0, 100, 256):
        one, _ = simple_box_model(cuda)
        one.set(**{k: float(v[m]) for k, v in ics.items()})
        for _ in range(5):
            one.time_step(20 * minutes)
        for name in ens.fields:
            assert ens.fields[name].interior.reshape(-1)[m].item() == one.fields[name].interior.item(), (m, name)


@pytest.mark.parametrize("timestepper", ["RungeKutta3", "Euler"])
def test_graph_replay_is_bit_identical_to_eager(cuda, timestepper):
    n, steps = 4096, 30
    rng = np.random.default_rng(4)
    ics = {k: torch.from_numpy(v * rng.uniform(0.5, 1.5, n)) for k, v in DEFAULTS.items()}
    T_fn = lambda t: 12.0 + 3.0 * math.sin(2 * math.pi * t / day)  # noqa: E731
    src = lambda t: 1e-7 * (1 + math.cos(2 * math.pi * t / day))  # noqa: E731

    def build():
        grid = ob.BoxModelGrid(n, device=cuda)
        PAR = ob.CenterField(grid, "PAR")
        bgc = ob.NPZD(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR)) if False else \
            ob.LOBSTER(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR), scale_negatives=True)
        m = ob.BoxModel(biogeochemistry=bgc, grid=grid, timestepper=timestepper, forcing={"NO₃": src},
                        prescribed_tracers={"PAR": PAR_fn, "T": T_fn})
        m.set(**ics)
        return m

    eager, graph = build(), build()
    oe = eager.run(20 * minutes, steps, graph=False, output_every=10)
    og_ = graph.run(20 * minutes, steps, graph=True, output_every=10)
    torch.cuda.synchronize()
    for name in eager.prognostic:
        assert torch.equal(eager.fields[name].data, graph.fields[name].data), name
        assert torch.equal(oe[name], og_[name]) and oe[name].shape == (3, n)
    assert graph.clock.iteration == steps and abs(graph.clock.time - eager.clock.time) < 1e-6
    assert not torch.equal(oe["P"][0], oe["P"][-1])  # something is happening (test_boxmodel.jl:81)


def test_box_nitrogen_is_conserved(cuda):
    model, _ = simple_box_model(cuda, 64)
    rng = np.random.default_rng(5)
    model.set(**{k: torch.from_numpy(v * rng.uniform(0.5, 1.5, 64)) for k, v in DEFAULTS.items()},
              sPOM=0.1, bPOM=0.1, DOM=0.2)
    N = lambda: sum(model.fields[k].interior.reshape(-1) for k in ("NO₃", "NH₄", "P", "Z", "sPOM", "bPOM", "DOM"))  # noqa: E731
    n0 = N().clone()
    model.run(20 * minutes, 100)
    assert torch.max(torch.abs(N() - n0) / n0).item() < 1e-13 * 100


def test_graph_mode_refuses_host_evaluated_time_dependence(cuda):
    grid = ob.BoxModelGrid(4, device=cuda)
    bgc = ob.LOBSTER(grid)  # default surface PAR is a function of time evaluated on the host
    m = ob.BoxModel(biogeochemistry=bgc, grid=grid)
    m.set(**DEFAULTS)
    with pytest.raises(ValueError):
        m.run(60.0, 3, graph=True)


def test_reference_simulation_with_speedy_output(cuda, tmp_path):
    """test/test_boxmodel.jl:55-79: 1000 steps of 20 minutes, `SpeedyOutput` every 20 iterations — every model field
    plus `t` is written, 1000/20 + 1 entries each (iteration 0 included), and every series moves."""
    model, _ = simple_box_model(cuda)
    model.set(**DEFAULTS)
    fname = str(tmp_path / "box_model_test.jld2")  # the reference's file name; the content is .npz (see SpeedyOutput)
    fast_output = ob.SpeedyOutput(fname)
    model.run(dt=20 * minutes, steps=1000, output_every=20, output=fast_output)
    results = ob.load_output(fast_output)
    assert len(results) == len(model.fields) + 1
    assert len(results["t"]) == 1000 // 20 + 1
    assert results["t"][0] == 0.0 and math.isclose(results["t"][-1], 1000 * 20 * minutes, rel_tol=1e-12)
    assert all(r[0].tolist() != r[-1].tolist() for r in results.values())
    assert np.array_equal(ob.load_output(fast_output, "P"), results["P"])


# ---- f-2: tendencies + tracer update in one launch (obm_npd_tendencies_substep) ---------------------------------------
@pytest.mark.parametrize("timestepper,graph", [("RungeKutta3", False), ("Euler", False), ("RungeKutta3", True)])
def test_fused_stage_is_bit_identical_to_the_three_launch_path(cuda, timestepper, graph):
    """`BoxModel(fused_step=True)`: compute_tendencies! + rk3_substep! + cache_previous_tendencies! (timesteppers.jl:30-93)
    as ONE launch per stage.  LOBSTER + carbonates + O₂ with a forcing and two prescribed series, eager and as a replayed
    CUDA graph: every tracer and every G⁻ equal those of the three-launch path bit for bit."""
    n, steps = 1500, 12
    rng = np.random.default_rng(14)
    ics = {k: torch.from_numpy(v * rng.uniform(0.5, 1.5, n)) for k, v in DEFAULTS.items()}
    ics.update(sPOM=0.2, bPOM=0.1, DOM=0.3, DIC=2200.0, Alk=2400.0, **{"O₂": 240.0})
    T_fn = lambda t: 12.0 + 3.0 * math.sin(2 * math.pi * t / day)  # noqa: E731
    src = lambda t: 1e-7 * (1 + math.cos(2 * math.pi * t / day))  # noqa: E731

    def build(fused):
        grid = ob.BoxModelGrid(n, device=cuda)
        PAR = ob.CenterField(grid, "PAR")
        bgc = ob.LOBSTER(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR),
                         carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen())
        m = ob.BoxModel(biogeochemistry=bgc, grid=grid, timestepper=timestepper, forcing={"NO₃": src},
                        prescribed_tracers={"PAR": PAR_fn, "T": T_fn}, fused_step=fused)
        m.set(**ics)
        return m

    three, one = build(False), build(True)
    three.run(20 * minutes, steps, graph=graph)
    one.run(20 * minutes, steps, graph=graph)
    torch.cuda.synchronize()
    for name in three.prognostic:
        assert torch.equal(three.fields[name].data, one.fields[name].data), name
        assert torch.equal(three.Gm[name].data, one.Gm[name].data), name
    assert not torch.equal(one.fields["P"].interior.reshape(-1), ics["P"].to(cuda))
    assert one.clock.iteration == steps and abs(one.clock.time - three.clock.time) < 1e-9


def test_fused_stage_against_the_oracle_and_argument_errors(cuda, oracle):
    """One fused launch on a 3-D grid against the oracle's tendencies followed by the oracle's substep (bit-exact update
    arithmetic ⇒ the tracers differ only through the tendencies' 1e-12); Gⁿ stored on request; T and tracers without a G⁻
    are left alone; refused: a stepped field that is not the tracer, accumulate without Gⁿ."""
    import ctypes as C
    from oceanbiome_b200 import _lib, synthetic
    grid = ob.RectilinearGrid(size=(40, 3, 5), extent=(1, 1, 50), device=cuda)
    og = oracle.Grid.like(grid)
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen()).underlying_biogeochemistry
    names = list(bgc.required_biogeochemical_tracers())
    rng = np.random.default_rng(21)
    hU = {m: rng.uniform(0.05, 2.0, og.parent_shape) for m in names}
    hGm = {m: rng.normal(size=og.parent_shape) * 1e-6 for m in names}
    hPAR = rng.uniform(0.0, 80.0, og.parent_shape)
    mk = lambda h: ob.CenterField(grid).set(og.interior(h))  # noqa: E731
    U, Gm, Gn = {m: mk(hU[m]) for m in names}, {m: mk(hGm[m]) for m in names}, {m: ob.CenterField(grid, fill=7.0) for m in names}
    for m in names:
        U[m].data.copy_(torch.from_numpy(hU[m])); Gm[m].data.copy_(torch.from_numpy(hGm[m]))
    PAR = ob.CenterField(grid); PAR.data.copy_(torch.from_numpy(hPAR))
    skip = "sPOM"                                       # no G⁻ given: not stepped
    dt, gamma, zeta = 900.0, 5 / 12, -17 / 60
    bgc.compute_tendencies_and_substep(grid, U, {"PAR": PAR}, {m: f for m, f in Gm.items() if m != skip}, dt, gamma, zeta,
                                       G=Gn, store_Gn=True)
    torch.cuda.synchronize()
    params = bgc.c_params()
    want_G = oracle.npd_tendencies(og, params, [hU[m] for m in names], hPAR)
    wU, wGm = [hU[m].copy() for m in names], [hGm[m].copy() for m in names]
    oracle.rk3_substep(og, wU, want_G, wGm, dt, gamma, zeta, cache_previous=True)
    for q, m in enumerate(names):
        gotU, gotGm, gotGn = (og.interior(host(f)) for f in (U[m], Gm[m], Gn[m]))
        if m == skip:
            assert np.array_equal(gotU, og.interior(hU[m])) and np.array_equal(gotGm, og.interior(hGm[m]))
        else:
            scale = np.maximum(np.abs(og.interior(wU[q])), dt * np.abs(og.interior(want_G[q])))
            assert np.max(np.abs(gotU - og.interior(wU[q])) / scale) <= 10 * RTOL_TENDENCY, m
            assert np.array_equal(gotGm, gotGn), m       # G⁻ ← Gⁿ, the value just computed
        g = og.interior(want_G[q])
        assert np.max(np.abs(gotGn - g)) <= 10 * RTOL_TENDENCY * max(np.max(np.abs(g)), 1e-30), m
        full = host(U[m]).copy(); og.interior(full)[...] = og.interior(hU[m])
        assert np.array_equal(full, hU[m]), m            # halos untouched
    # argument errors
    lib = _lib.load()
    cg = grid.c_grid()
    tab = lambda fs: _lib.pointer_table([f.ptr for f in fs])  # noqa: E731
    order = [U[m] for m in names]
    other = [ob.CenterField(grid) for _ in names]
    p = bgc.c_params()
    call = lambda t, G, acc, store, m: lib.obm_npd_tendencies_substep(  # noqa: E731
        C.byref(cg), C.byref(p), 0, None, None, t, PAR.ptr, G, acc, store, m, dt, gamma, zeta, 1, None)
    assert call(tab(order), None, 1, 0, tab([Gm[m] for m in names])) == -1     # accumulate needs Gⁿ
    assert call(tab(order), None, 0, 0, None) == -1                            # no G⁻ table
    assert call(None, None, 0, 0, tab(other)) == -1


# ---- f-3: the whole run in one launch (obm_npd_box_run) ----------------------------------------------------------------
@pytest.mark.parametrize("timestepper,model", [("RungeKutta3", "lobster"), ("Euler", "lobster"), ("RungeKutta3", "npzd_sweep")])
def test_whole_run_launch_is_bit_identical_to_the_replayed_graph(cuda, timestepper, model):
    """`run(..., device_loop=True)`: every thread integrates its box through all stages of all steps in ONE launch, the
    prescribed series read from the same device tables.  Tracers, G⁻, the PAR / T fields left behind and every snapshot
    equal those of the replayed CUDA graph of per-stage launches bit for bit — LOBSTER + carbonates + O₂ with prescribed
    PAR and temperature over a ragged number of boxes, and the reference's NPZD box benchmark as a parameter sweep with a
    per-box PAR series; argument errors are refused."""
    n, steps, every = (1500, 12, 4) if model == "lobster" else (333, 9, 3)
    rng = np.random.default_rng(5)
    T_fn = lambda t: 12.0 + 3.0 * math.sin(2 * math.pi * t / day)  # noqa: E731
    if model == "lobster":
        ics = {k: torch.from_numpy(v * rng.uniform(0.5, 1.5, n)) for k, v in DEFAULTS.items()}
        ics.update(sPOM=0.2, bPOM=0.1, DOM=0.3, DIC=2200.0, Alk=2400.0, **{"O₂": 240.0})
        par = PAR_fn
    else:
        ics = {"N": torch.from_numpy(rng.uniform(5.0, 12.0, n)), "P": 0.1, "Z": 0.01, "D": 0.0}
        scale = torch.from_numpy(rng.uniform(0.5, 1.5, n))
        par = lambda t: PAR_fn(t) * scale  # noqa: E731  (one PAR series per box)

    def build():
        grid = ob.BoxModelGrid(n, device=cuda)
        PAR = ob.CenterField(grid, "PAR")
        light = ob.PrescribedPhotosyntheticallyActiveRadiation(PAR)
        if model == "lobster":
            bgc = ob.LOBSTER(grid, light_attenuation=light, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen())
        else:
            sweep = {"phytoplankton_maximum_growth_rate": torch.from_numpy(np.linspace(0.5, 1.5, n) * 0.6989 / day),
                     "maximum_grazing_rate": torch.from_numpy(np.linspace(1.5, 0.5, n) * 2.1522 / day)}
            bgc = ob.NPZD(grid, light_attenuation=light, parameter_ensemble=sweep)
        m = ob.BoxModel(biogeochemistry=bgc, grid=grid, timestepper=timestepper, prescribed_tracers={"PAR": par, "T": T_fn},
                        fused_step=True)
        m.set(**ics)
        return m

    a, b = build(), build()
    ra = a.run(20 * minutes, steps, graph=True, output_every=every)
    rb = b.run(20 * minutes, steps, device_loop=True, output_every=every)
    torch.cuda.synchronize()
    for name in a.prognostic:
        assert torch.equal(a.fields[name].data, b.fields[name].data), name
        assert torch.equal(a.Gm[name].data, b.Gm[name].data), name
        assert torch.equal(ra[name], rb[name]) and ra[name].shape == (steps // every, n), name
    assert torch.equal(a.auxiliary_fields["PAR"].data, b.auxiliary_fields["PAR"].data)
    assert torch.equal(a.fields["T"].data, b.fields["T"].data)
    assert b.clock.iteration == steps and abs(a.clock.time - b.clock.time) < 1e-9
    assert not torch.equal(b.fields["P"].interior.reshape(-1), torch.as_tensor(ics["P"], dtype=torch.float64).expand(n).to(cuda))
    # a second call continues from where the first one stopped, like a second replay loop
    a.run(20 * minutes, 3, graph=True)
    b.run(20 * minutes, 3, device_loop=True)
    torch.cuda.synchronize()
    for name in a.prognostic:
        assert torch.equal(a.fields[name].data, b.fields[name].data), name
    # refused: no fused step; a forcing; a prescribed biogeochemical tracer
    grid = ob.BoxModelGrid(4, device=cuda)
    PAR = ob.CenterField(grid, "PAR")
    bgc = ob.NPZD(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR))
    with pytest.raises(ValueError, match="fused_step"):
        ob.BoxModel(biogeochemistry=bgc, grid=grid, prescribed_tracers={"PAR": PAR_fn, "T": T_fn}).run(60.0, 2, device_loop=True)
    with pytest.raises(ValueError, match="forcings"):
        ob.BoxModel(biogeochemistry=bgc, grid=grid, prescribed_tracers={"PAR": PAR_fn, "T": T_fn}, forcing={"N": lambda t: 1e-9},
                    fused_step=True).run(60.0, 2, device_loop=True)
    with pytest.raises(ValueError, match="prescribes PAR"):
        ob.BoxModel(biogeochemistry=bgc, grid=grid, prescribed_tracers={"PAR": PAR_fn, "T": T_fn, "Z": lambda t: 0.05},
                    fused_step=True).run(60.0, 2, device_loop=True)
