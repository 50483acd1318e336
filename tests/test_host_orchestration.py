"""CPU: which C entry points the hooks call, and in which order — the reference's fixed order
`update_biogeochemical_state!`: modifiers → light → underlying → sediment (src/OceanBioME.jl:161-167) and
`update_tendencies!`: underlying → sediment → particles (:148-152) — for every fusion the host layer can choose.
The library is replaced by a recorder (no kernel runs, no GPU needed); argument COUNTS are checked against the ctypes
prototypes so a drifted call site fails here rather than on the GPU box."""
import importlib

import pytest

import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib

MODULES = ["biogeochemistry", "negative_tracers", "light", "pisces", "carbon_chemistry", "sediments", "npd", "particles",
           "gas_exchange", "box_model"]


class Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name not in _lib.PROTOTYPES:
            raise AttributeError(name)
        nargs = len(_lib.PROTOTYPES[name][1])

        def fn(*args):
            assert len(args) == nargs, f"{name}: {len(args)} arguments, prototype has {nargs}"
            self.calls.append(name)
            return 0
        return fn


@pytest.fixture
def lib(monkeypatch):
    rec = Recorder()
    monkeypatch.setattr(_lib, "load", lambda path=None: rec)
    for m in MODULES:
        mod = importlib.import_module(f"oceanbiome_b200.{m}")
        if hasattr(mod, "require_cuda"):
            monkeypatch.setattr(mod, "require_cuda", lambda *a, **k: None)
        if hasattr(mod, "current_stream_ptr"):
            monkeypatch.setattr(mod, "current_stream_ptr", lambda device: None)
    return rec


def grid3():
    return ob.RectilinearGrid(size=(4, 3, 8), extent=(4.0, 3.0, 80.0), device="cpu")


def test_pisces_stage_is_three_launches(lib):
    g = grid3()
    bgc = ob.PISCES(g, scale_negatives=True)
    model = ob.BiogeochemicalModel(g, bgc)
    model.update_state()
    assert lib.calls == ["obm_scale_negative_tracers_calcite_saturation", "obm_par_multiband_column_state"]
    lib.calls.clear()
    bgc.update_tendencies(model)
    assert lib.calls == ["obm_pisces_tendencies"]


def test_pisces_unfused_follows_the_reference_launch_by_launch(lib):
    g = grid3()
    bgc = ob.PISCES(g, scale_negatives=True)
    bgc.fuse_state_update = False
    ob.BiogeochemicalModel(g, bgc).update_state()
    assert lib.calls == ["obm_scale_negative_tracers", "obm_par_multiband", "obm_euphotic_depth", "obm_mixed_layer_mean",
                         "obm_calcite_saturation"]


def test_pisces_without_modifiers_keeps_the_omega_launch(lib):
    g = grid3()
    ob.BiogeochemicalModel(g, ob.PISCES(g)).update_state()
    assert lib.calls == ["obm_par_multiband_column_state", "obm_calcite_saturation"]


def test_pisces_with_prescribed_light_uses_the_standalone_column_kernels(lib):
    g = grid3()
    PAR = {n: ob.CenterField(g, n, 10.0) for n in ("PAR₁", "PAR₂", "PAR₃", "PAR")}
    bgc = ob.PISCES(g, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR), scale_negatives=True)
    ob.BiogeochemicalModel(g, bgc).update_state()
    assert lib.calls == ["obm_scale_negative_tracers_calcite_saturation", "obm_euphotic_depth", "obm_mixed_layer_mean"]


def test_a_modifier_after_the_scalers_blocks_the_omega_fusion(lib):
    g = grid3()
    bgc = ob.PISCES(g, scale_negatives=True)
    bgc.modifiers = (*bgc.modifiers, ob.ZeroNegativeTracers())  # the last launch is no longer a negative scaling
    ob.BiogeochemicalModel(g, bgc).update_state()
    assert lib.calls == ["obm_scale_negative_tracers", "obm_zero_negative_tracers", "obm_par_multiband_column_state",
                         "obm_calcite_saturation"]


def test_lobster_with_sediment_order(lib):
    g = grid3()
    sed = ob.SimpleMultiGSediment(g)
    bgc = ob.LOBSTER(g, oxygen=ob.Oxygen(), sediment=sed, scale_negatives=True)
    model = ob.BiogeochemicalModel(g, bgc)
    model.update_state()
    assert lib.calls == ["obm_scale_negative_tracers", "obm_par_twoband", "obm_sediment_update_state"]
    lib.calls.clear()
    model.compute_tendencies()
    assert lib.calls == ["obm_npd_tendencies", "obm_sediment_update_tendencies"]


def test_particles_and_sinking_in_a_time_step(lib):
    g = grid3()
    kelp = ob.SugarKelpParticles(5, g, coupled_tracers={"NO₃": "NO₃", "NH₄": "NH₄", "DON": "DOM", "bPON": "bPOM"})
    bgc = ob.LOBSTER(g, particles=kelp)
    model = ob.BiogeochemicalModel(g, bgc, extra_tracers=("T",), timestepper="Euler", sinking_advection="UpwindBiased1")
    model.time_step(10.0)
    # first step: no particle step yet (Oceananigans steps particles at the END of a stage, with that stage's Δt)
    assert lib.calls == ["obm_par_twoband", "obm_npd_tendencies", "obm_kelp_update_tendencies", "obm_sinking_tendencies",
                         "obm_rk3_substep"]
    lib.calls.clear()
    model.time_step(10.0)
    assert lib.calls == ["obm_par_twoband", "obm_npd_tendencies", "obm_kelp_update_tendencies", "obm_sinking_tendencies",
                         "obm_kelp_step", "obm_rk3_substep"]
    lib.calls.clear()
    model.finish_particles()
    assert lib.calls[-1] == "obm_kelp_step" and model.clock.last_stage_dt == float("inf")


def test_rk3_model_calls_the_hooks_once_per_stage(lib):
    g = grid3()
    model = ob.BiogeochemicalModel(g, ob.NPZD(g))
    model.time_step(60.0)
    assert lib.calls == ["obm_par_twoband", "obm_npd_tendencies", "obm_rk3_substep"] * 3
    assert model.clock.iteration == 1 and abs(model.clock.time - 60.0) < 1e-9


def test_box_model_follows_oceananigans_rk3_order(lib):
    """boxmodel.jl:92-110 + Oceananigans' RK3 `time_step!`: state and tendencies at iteration 0, then per stage
    substep → update_state! (state hooks, then tendencies)."""
    g = ob.BoxModelGrid(3, device="cpu")
    PAR = ob.CenterField(g, "PAR")
    bgc = ob.LOBSTER(g, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR))
    box = ob.BoxModel(biogeochemistry=bgc, grid=g, prescribed_tracers={"PAR": lambda t: 50.0})
    box.set(P=0.1, Z=0.01)
    box.time_step(1200.0)
    assert lib.calls == ["obm_npd_tendencies"] + ["obm_rk3_substep", "obm_npd_tendencies"] * 3
    lib.calls.clear()
    box.time_step(1200.0)  # no initial update the second time
    assert lib.calls == ["obm_rk3_substep", "obm_npd_tendencies"] * 3
    assert float(PAR.interior[0, 0, 0]) == 50.0


def test_gas_exchange_boundary_condition_is_applied_after_the_tendencies(lib):
    g = grid3()
    co2 = ob.CarbonDioxideGasExchangeBoundaryCondition(air_concentration=413.0, wind_speed=2.0)
    bgc = ob.LOBSTER(g, carbonate_system=ob.CarbonateSystem())
    model = ob.BiogeochemicalModel(g, bgc, extra_tracers=("T", "S"), boundary_conditions={"DIC": co2})
    model.compute_tendencies()
    assert lib.calls == ["obm_npd_tendencies", "obm_gas_exchange_flux"]


def test_parameter_ensemble_routes_to_the_ensemble_entry_point(lib):
    g = grid3()
    growth = [1e-5 * (1 + m) for m in range(g.Nx * g.Ny)]
    bgc = ob.NPZD(g, parameter_ensemble={"phytoplankton_maximum_growth_rate": growth})
    model = ob.BiogeochemicalModel(g, bgc)
    bgc.update_tendencies(model)
    assert lib.calls == ["obm_npd_tendencies_ensemble"]
    lib.calls.clear()
    bgc.underlying_biogeochemistry.set_parameter_ensemble()  # back to one parameter set
    bgc.update_tendencies(model)
    assert lib.calls == ["obm_npd_tendencies"]


def test_parameter_ensemble_checks_names_and_member_count(lib):
    g = grid3()
    u = ob.NPZD(g).underlying_biogeochemistry
    with pytest.raises(KeyError):
        u.set_parameter_ensemble(no_such_rate=[1.0] * 12)
    with pytest.raises(ValueError):
        u.set_parameter_ensemble(maximum_grazing_rate=[1.0] * 12, grazing_half_saturation=[1.0] * 11)
    bgc = ob.NPZD(g, parameter_ensemble={"maximum_grazing_rate": [1.0] * 5})
    with pytest.raises(ValueError, match="5 members"):
        bgc.update_tendencies(ob.BiogeochemicalModel(g, bgc))


def pisces_box(constant_zeu=True, constant_mean_light=True):
    from oceanbiome_b200 import pisces
    grid = ob.BoxModelGrid(1, device="cpu", z=-5)
    PAR = {n: ob.CenterField(grid, n, 100.0) for n in ("PAR₁", "PAR₂", "PAR₃")}
    PAR["PAR"] = ob.CenterField(grid, "PAR", 300.0)
    bgc = ob.PISCES(grid, sinking_speeds={"POC": 0.0, "GOC": 0.0},
                    light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR),
                    mixed_layer_depth=ob.ConstantField(grid, -10.0),
                    euphotic_depth=ob.ConstantField(grid, -10.0) if constant_zeu else None,
                    mean_mixed_layer_vertical_diffusivity=ob.ConstantField(grid, 1.0),
                    mean_mixed_layer_light=ob.ConstantField(grid, 300.0) if constant_mean_light else None,
                    iron=pisces.SimpleIron(excess_scavenging_enhancement=0.0),
                    nitrogen=pisces.NitrateAmmonia(maximum_fixation_rate=0.0))
    return ob.BoxModel(biogeochemistry=bgc, grid=grid), bgc


def test_pisces_box_model_leaves_constant_fields_alone(lib):
    """test/test_PISCES.jl:32-60: a PISCES box model prescribes zₑᵤ, κ̄ and PAR̄ₘₓₗ as ConstantFields, for which the
    reference's state update does nothing (compute_euphotic_depth.jl:44-46, mean_mixed_layer_properties.jl:21,59)."""
    model, bgc = pisces_box()
    model.time_step(1.0)
    stage = ["obm_calcite_saturation", "obm_pisces_tendencies"]
    assert lib.calls == stage + 3 * (["obm_rk3_substep"] + stage)
    u = bgc.underlying_biogeochemistry
    assert u.euphotic_depth.data.unique().tolist() == [-10.0] and u.mean_mixed_layer_light.data.unique().tolist() == [300.0]


def test_pisces_recomputes_only_what_is_not_constant(lib):
    model, _ = pisces_box(constant_zeu=False)
    model.update_state()
    assert lib.calls == ["obm_euphotic_depth", "obm_calcite_saturation", "obm_pisces_tendencies"]
    lib.calls.clear()
    model, _ = pisces_box(constant_mean_light=False)
    model.update_state()
    assert lib.calls == ["obm_mixed_layer_mean", "obm_calcite_saturation", "obm_pisces_tendencies"]


def test_computed_light_with_a_constant_column_field_takes_the_plain_scan(lib):
    g = grid3()
    bgc = ob.PISCES(g, euphotic_depth=ob.ConstantField(g, -50.0))
    ob.BiogeochemicalModel(g, bgc).update_state()
    assert lib.calls == ["obm_par_multiband", "obm_mixed_layer_mean", "obm_calcite_saturation"]


def test_pisces_per_tracer_call_form_is_one_fused_launch(lib):
    """`bgc(i, j, k, grid, Val(name), clock, fields, auxiliary_fields)` (PISCES.jl:120-123) in the host mirror: one
    `obm_pisces_tendencies` launch on a row of boxes, whatever the tracer asked for; unknown names are refused."""
    import torch
    g = grid3()
    u = ob.PISCES(g).underlying_biogeochemistry
    out = u("P", device="cpu", P=1.0, PChl=0.3, PFe=0.01, **{"NO₃": 4.0, "PAR₁": 20.0, "PAR₂": 20.0, "PAR₃": 20.0, "zₘₓₗ": -30.0})
    assert isinstance(out, float) and lib.calls == ["obm_pisces_tendencies"]
    lib.calls.clear()
    many = u("Fe", device="cpu", z=-120.0, time=86400.0, Fe=torch.linspace(0.1, 1.0, 5, dtype=torch.float64), **{"O₂": 200.0})
    assert many.shape == (5,) and lib.calls == ["obm_pisces_tendencies"]
    with pytest.raises(KeyError):
        u("Q", device="cpu")
    with pytest.raises(KeyError):
        u("P", device="cpu", Q=1.0)


def test_model_latitude_goes_through_the_per_row_entry_point(lib):
    """`latitude = ModelLatitude()` (PISCES/common.jl:27-28) on a grid with its own latitude: same hooks, same order, the
    tendencies through `obm_pisces_tendencies_rows` with a [3][Ny] table; a prescribed latitude keeps the scalar entry point."""
    g = ob.LatitudeLongitudeGrid(size=(4, 6, 8), longitude=(0, 4), latitude=(-30, 30), z=(-80, 0), device="cpu")
    bgc = ob.PISCES(g, latitude=ob.ModelLatitude(), scale_negatives=True)
    model = ob.BiogeochemicalModel(g, bgc)
    model.update_state()
    assert lib.calls == ["obm_scale_negative_tracers_calcite_saturation", "obm_par_multiband_column_state"]
    lib.calls.clear()
    bgc.update_tendencies(model)
    assert lib.calls == ["obm_pisces_tendencies_rows"]
    assert tuple(bgc.underlying_biogeochemistry._row_table.shape) == (3, 6)
    lib.calls.clear()
    bgc2 = ob.PISCES(grid3(), latitude=ob.PrescribedLatitude(-12.0))
    m2 = ob.BiogeochemicalModel(bgc2.underlying_biogeochemistry.grid, bgc2)
    bgc2.update_tendencies(m2)
    assert lib.calls == ["obm_pisces_tendencies"]


def test_whole_run_of_a_box_ensemble_is_one_call(lib):
    """`BoxModel.run(device_loop=True)`: the tabulated series go to ONE `obm_npd_box_run` whatever the number of steps; the
    clock ends where the per-stage path would leave it; a prescribed series nothing reads still ends up in its field."""
    import math
    grid = ob.BoxModelGrid(5, device="cpu")
    PAR = ob.CenterField(grid, "PAR")
    bgc = ob.LOBSTER(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR))
    m = ob.BoxModel(biogeochemistry=bgc, grid=grid, prescribed_tracers={"PAR": lambda t: 10 + math.sin(t), "T": lambda t: 3.0 + t},
                    fused_step=True)
    m.set(**{"NO₃": 10.0, "NH₄": 0.1, "P": 0.1, "Z": 0.01})
    lib.calls.clear()
    out = m.run(600.0, 7, device_loop=True, output_every=2)
    assert lib.calls == ["obm_npd_box_run"]
    assert m.clock.iteration == 7 and abs(m.clock.time - 7 * 600.0) < 1e-6
    assert set(out) == set(m.prognostic) and out["P"].shape == (3, 5)
    assert abs(float(m.fields["T"].interior.reshape(-1)[0]) - (3.0 + m.clock.time)) < 1e-9  # LOBSTER does not read T


def test_column_model_run_loops_time_step_and_takes_snapshots(lib):
    """`BiogeochemicalModel.run` without a graph is the eager loop of `time_step` (three stages of hooks + one substep launch each)
    with interior snapshots every `output_every` steps; PISCES is refused by the graph mode before anything is captured."""
    import pytest
    g = grid3()
    model = ob.BiogeochemicalModel(g, ob.LOBSTER(g, scale_negatives=True), sinking_advection="UpwindBiased1")
    lib.calls.clear()
    out = model.run(60.0, 4, output_every=2)
    stage = ["obm_scale_negative_tracers", "obm_par_twoband", "obm_npd_tendencies", "obm_sinking_tendencies", "obm_rk3_substep"]
    assert lib.calls == stage * 12
    assert out["P"].shape == (2, g.Nz, g.Ny, g.Nx) and model.clock.iteration == 4 and abs(model.clock.time - 240.0) < 1e-9
    with pytest.raises(ValueError, match="not PISCES"):
        ob.BiogeochemicalModel(g, ob.PISCES(g)).run(60.0, 5, graph=True)
