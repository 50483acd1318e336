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
