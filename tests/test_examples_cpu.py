"""CPU: the host-side pieces of examples/data_assimilation.py (the Kalman update, the priors, the observables) — the
forward model itself needs the GPU (tests/test_gpu_examples.py)."""
import importlib.util
import math
import os

import numpy as np
import torch


def load_example():
    spec = importlib.util.spec_from_file_location(
        "data_assimilation", os.path.join(os.path.dirname(__file__), "..", "examples", "data_assimilation.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_ensemble_kalman_inversion_recovers_a_linear_model():
    da, rng = load_example(), np.random.default_rng(0)
    A, truth, Gamma = rng.normal(size=(5, 4)), np.array([1.0, 2.0, 3.0, 4.0]), np.eye(5) * 1e-4
    y = A @ truth
    theta = rng.normal(size=(4, 40)) * 2 + 2
    for _ in range(10):
        theta = da.update_ensemble(theta, A @ theta, y, Gamma, rng)
    assert np.allclose(theta.mean(axis=1), truth, atol=0.02)
    # a failed member (NaN observables) is redrawn from the successful ones, not propagated
    g = A @ theta
    g[:, 3] = np.nan
    new = da.update_ensemble(theta, g, y, Gamma, rng)
    assert new.shape == theta.shape and np.isfinite(new).all()


def test_constrained_gaussian_has_the_requested_moments():
    da, rng = load_example(), np.random.default_rng(1)
    x = np.exp(da.constrained_gaussian(rng, 0.7, 0.1, 200_000))
    assert (x > 0).all() and abs(x.mean() - 0.7) < 2e-3 and abs(x.std() - 0.1) < 2e-3


def test_observables_of_a_known_series():
    da = load_example()
    day = da.day
    n = 1093
    times = (np.arange(n) + 1) * 8 * 3600.0
    t = torch.as_tensor(times)
    # member 0: a bloom peaking at output 300 and collapsing fastest right after output 700; member 1 goes negative
    P0 = 0.1 + torch.exp(-((t - times[300]) / (20 * day)) ** 2) + 0.5 / (1 + torch.exp((t - times[700]) / (2 * day)))
    P1 = P0 - 1.0
    obs = da.extract_observables(torch.stack([P0, P1], dim=1), times)
    assert obs.shape == (5, 2) and np.isnan(obs[:, 1]).all()  # "model failure" ⇒ NaN observables (data_assimilation.jl:98)
    peak, winter, average, peak_timing, die_off = obs[:, 0]
    assert math.isclose(peak, float(P0.max())) and math.isclose(winter, float(P0.min())) and math.isclose(average, float(P0.mean()))
    assert abs(peak_timing - times[300] / day) < 1.0 and abs(die_off - times[700] / day) < 1.0


def test_parameter_mapping_of_the_example():
    da = load_example()
    u = np.array([[0.15 / da.day], [0.7 / da.day], [2.4], [0.01 / da.day]])
    p = da.phytozoo_parameters(u)
    assert math.isclose(p["light_half_saturation"][0], 0.7 / 0.15)
    # as the reference writes it (data_assimilation.jl:42): the base mortality, already per second, is divided by `day` again
    assert math.isclose(p["phytoplankton_mortality_rate"][0], 0.066 / da.day + (0.01 / da.day) / da.day)
    assert math.isclose(p["phytoplankton_solid_waste_fraction"][0], 0.01 / 0.076)
    # every key is a parameter the ensemble entry point can vary
    import oceanbiome_b200 as ob
    for k in p:
        assert ob.NutrientsPlanktonDetritus.parameter_index(k) >= 0


def test_box_example_builds_the_reference_setup():
    """examples/box.py (the reference's examples/box.jl): construction, initial values and the PAR cycle; the run itself
    is the graph-replay path of tests/test_gpu_box_model.py."""
    spec = importlib.util.spec_from_file_location("box", os.path.join(os.path.dirname(__file__), "..", "examples", "box.py"))
    box = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(box)
    model = box.build(3, device="cpu")
    assert model.prognostic == ["NO₃", "NH₄", "P", "Z", "sPOM", "bPOM", "DOM"] and model.grid.Nx == 3
    assert model.fields["NO₃"].interior.unique().tolist() == [10.0] and model.fields["Z"].interior.unique().tolist() == [0.01]
    # box.jl:24-28: PAR⁰ between 2 and ≈ 122 W m⁻², attenuated by exp(0.2 · (−10 m))
    vals = [box.PAR_func(d * box.day) for d in range(0, 365, 5)]
    assert min(vals) >= 2 * math.exp(-2) and max(vals) <= 122 * math.exp(-2) and vals.index(max(vals)) * 5 in range(140, 200)
    assert math.isclose(box.PAR_func(0.0), box.PAR_func(box.year))  # one-year period
