"""GPU: the usage examples of INTEGRATION.md run as written (small sizes)."""
import pytest
import torch

import oceanbiome_b200 as ob

pytestmark = pytest.mark.gpu


def test_column_model_with_sediment_kelp_and_sinking(cuda):
    col = ob.RectilinearGrid(size=(64, 1, 32), extent=(64.0, 1.0, 200.0), device=cuda)
    sed = ob.SimpleMultiGSediment(col)
    kelp = ob.SugarKelpParticles(100, col, coupled_tracers={"NO₃": "NO₃", "NH₄": "NH₄", "DON": "DOM", "bPON": "bPOM"})
    bgc = ob.LOBSTER(col, oxygen=ob.Oxygen(), sediment=sed, particles=kelp, scale_negatives=True)
    model = ob.BiogeochemicalModel(col, bgc, extra_tracers=("T",), sinking_advection="UpwindBiased3")
    model.set(P=0.1, Z=0.05, sPOM=0.2, bPOM=0.2, T=10.0, **{"NO₃": 5.0, "NH₄": 0.2, "O₂": 250.0})
    for f in model.tracers.values():
        f.fill_halos_zero_gradient()
    kelp.set(x=32.0, z=-5.0, A=10.0, N=0.015, C=0.3)
    model.clock.time = 100 * 86400.0
    A0 = kelp.fields["A"].clone()
    for _ in range(3):
        model.time_step(600.0)
    model.finish_particles()
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(f.interior).all()) for f in model.tracers.values())
    assert not torch.equal(kelp.fields["A"], A0)                                  # the kelp grew (or eroded)
    assert all(bool((f.interior > 0).all()) for f in sed.fields.values())         # particles reached the sediment
    assert bool((model.tracers["NO₃"].interior[-2, 0, 32] < 5.0).item())          # uptake by the kelp's cell (z = −5 m)


def test_data_assimilation_example_shortened(cuda):
    """examples/data_assimilation.py (the reference's examples/data_assimilation.jl with the ensemble on the device):
    40 model days instead of 2 years, 2 iterations; the ensemble run equals member-by-member runs bit for bit and the
    Kalman update returns a finite ensemble."""
    import importlib.util
    import os
    import numpy as np
    spec = importlib.util.spec_from_file_location(
        "data_assimilation", os.path.join(os.path.dirname(__file__), "..", "examples", "data_assimilation.py"))
    da = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(da)
    day = 86400.0
    u = np.array([[0.15, 0.2, 0.25], [0.7, 0.6, 0.8], [2.4, 2.0, 3.0], [0.01, 0.02, 0.005]]) / np.array([[day], [day], [1], [day]])
    kw = dict(stop_time=40 * day, device=cuda)
    P, times = da.run_box_simulations(u, **kw)
    assert P.shape[1] == 3 and len(times) == P.shape[0] and bool(torch.isfinite(P).all())
    for m in range(3):
        Pm, _ = da.run_box_simulations(u[:, m:m + 1], **kw)
        assert torch.equal(Pm[:, 0], P[:, m])
    obs = da.extract_observables(P, times)
    assert obs.shape == (5, 3) and np.isfinite(obs).all() and len(set(obs[0])) == 3
    truth, final, history = da.main(N_ensemble=6, N_iterations=2, **kw)
    assert final.shape == (4, 6) and np.isfinite(final).all() and (final > 0).all() and len(history) == 2
