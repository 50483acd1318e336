"""GPU: the sugar-kelp particle kernels (obm_kelp_update_tendencies, obm_kelp_step) against the oracle on identical
inputs, and the reference's own test (test/test_sugar_kelp.jl) run through the hooks on the device."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu
day = 86400.0


def host(f):
    return np.ascontiguousarray(f.data.cpu().numpy())


class FakeModel:
    """the three things the particle hooks read from a model"""

    def __init__(self, grid, tracers, aux, Gn, time):
        self.grid, self.tracers, self.Gn = grid, tracers, Gn
        self.clock = ob.Clock(time=time)
        self.biogeochemistry = type("B", (), {"biogeochemical_auxiliary_fields": lambda self_: aux})()


@pytest.mark.parametrize("CN", [math.inf, 9.0])
def test_kernels_match_oracle(cuda, oracle, CN):
    grid = ob.RectilinearGrid(size=(9, 7, 11), x=(0, 18), y=(-7, 7), z=np.cumsum(np.r_[-40.0, 2.0 + np.arange(11) * 0.3]), device=cuda)
    og = oracle.Grid.like(grid)
    names = ("T", "NO₃", "NH₄", "PAR", "u")
    ranges = {"T": (4, 18), "NO₃": (0.5, 12), "NH₄": (0.05, 3), "PAR": (1, 150), "u": (0.0, 0.4)}
    fields = {n: synthetic.fill_torch(ob.CenterField(grid, n), n, *ranges[n]) for n in names}
    n = 300
    rng = np.random.default_rng(11)
    part = ob.BiogeochemicalParticles(n, grid, biogeochemistry=ob.SugarKelp(exudation_redfield_ratio=CN),
                                      scalefactors=rng.uniform(0.5, 3.0, n))
    part.set(x=rng.uniform(-3, 21, n), y=rng.uniform(-9, 9, n), z=rng.uniform(-45, 2, n),  # some outside: wrap / clamp
             A=rng.uniform(0.5, 40, n), N=rng.uniform(0.0127, 0.0215, n), C=rng.uniform(0.02, 0.6, n))
    coupled = part.biogeochemistry.coupled_tracers()
    Gn = {c: ob.CenterField(grid, "G" + c, 1e-9) for c in coupled if c != "DON"}  # DON not a tracer of this model
    t = 75 * day
    model = FakeModel(grid, {"T": fields["T"], "NO₃": fields["NO₃"], "NH₄": fields["NH₄"]}, {"PAR": fields["PAR"]}, Gn, t)
    model.velocities = {"u": fields["u"]}
    # oracle twins
    h = {k: host(v) for k, v in fields.items()}
    q0 = {k: v.cpu().numpy().copy() for k, v in part.fields.items()}
    qo, keep = oracle.make_particles(part.x.cpu().numpy(), part.y.cpu().numpy(), part.z.cpu().numpy(), q0["A"].copy(), q0["N"].copy(),
                                     q0["C"].copy(), part.scalefactors.cpu().numpy(), *[getattr(part.c_particles(), k) for k in ("x0", "dx", "y0", "dy")],
                                     tuple(part.c_particles().topology))
    fo, keep_f = oracle.make_kelp_tracers(h["T"], h["NO₃"], h["NH₄"], h["PAR"], u=h["u"])
    p = part.biogeochemistry.c_params()
    Go = [np.full(og.parent_shape, 1e-9) if c != "DON" else None for c in coupled]
    oracle.kelp_update_tendencies(og, p, qo, fo, Go, t)
    part.update_tendencies(None, model)
    torch.cuda.synchronize()
    for c, want in zip(coupled, Go):
        if want is None:
            continue
        got = host(Gn[c])
        touched = want != 1e-9
        assert touched.sum() > 50
        assert np.array_equal(got != 1e-9, touched) or c in ("NO₃",)  # a zero uptake adds exactly 0
        scale = np.abs(want - 1e-9).max()
        assert np.max(np.abs(got - want)) <= 1e-12 * scale, c
    # step
    out = [np.zeros(n) for _ in range(3)]
    oracle.kelp_step(og, p, qo, fo, t, 40.0, out)
    part.step(model, 40.0)
    torch.cuda.synchronize()
    for j, name in enumerate(("A", "N", "C")):
        d = part.tendencies[name].cpu().numpy()
        assert np.max(np.abs(d - out[j])) <= 1e-12 * np.abs(out[j]).max(), name
        np.testing.assert_allclose(part.fields[name].cpu().numpy(), keep[3 + j], rtol=1e-13)
        assert not np.array_equal(part.fields[name].cpu().numpy(), q0[name])


def test_argument_checks(cuda):
    grid = ob.RectilinearGrid(size=(2, 2, 2), extent=(1, 1, 1), device=cuda)
    lib = _lib.load()
    cg, p, q, f = grid.c_grid(), ob.SugarKelp().c_params(), _lib.obm_particles(), _lib.obm_kelp_tracers()
    q.n = 3
    assert lib.obm_kelp_step(C.byref(cg), C.byref(p), C.byref(q), C.byref(f), 0.0, 1.0, None, None) == -1  # NULL arrays
    q.n = -1
    assert lib.obm_kelp_step(C.byref(cg), C.byref(p), C.byref(q), C.byref(f), 0.0, 1.0, None, None) == -2
    q.n = 0
    assert lib.obm_kelp_step(C.byref(cg), C.byref(p), C.byref(q), C.byref(f), 0.0, 1.0, None, None) == 0
    with pytest.raises(NotImplementedError):
        ob.BiogeochemicalParticles(2, grid, biogeochemistry=object())


def test_reference_sugar_kelp_test_on_device(cuda):
    """test/test_sugar_kelp.jl:21-90: two kelp particles in a 1×1×1 box with LOBSTER (carbonates, variable Redfield,
    oxygen, no sinking), NO₃ = 10, NH₄ = 1, DIC = Alk = 2000, T = 10, S = 35, A = 2, N = C = 1 at t = 60 days, ten
    steps of Δt = 1: the kelp is integrated, and kelp + tracer nitrogen and carbon are conserved to the reference's
    CPU tolerance √eps(total)."""
    grid = ob.RectilinearGrid(size=(1, 1, 1), extent=(1, 1, 1), device=cuda)
    particles = ob.SugarKelpParticles(2, grid, advection=None)
    assert isinstance(particles, ob.BiogeochemicalParticles) and isinstance(particles.biogeochemistry, ob.SugarKelp)
    assert len(particles) == 2
    bgc = ob.LOBSTER(grid, particles=particles, carbonate_system=ob.CarbonateSystem(), detritus=ob.VariableRedfieldDetritus(),
                     oxygen=ob.Oxygen())
    model = ob.BiogeochemicalModel(grid, bgc, extra_tracers=("T", "S"))
    model.set(**{"NO₃": 10.0, "NH₄": 1.0, "DIC": 2000.0, "Alk": 2000.0, "T": 10.0, "S": 35.0})
    particles.set(x=0.5, y=0.5, z=-0.5, A=2.0, N=1.0, C=1.0)
    k = particles.biogeochemistry
    Ns, Cs, kA = k.structural_nitrogen, k.structural_carbon, k.structural_dry_weight_per_area
    u = bgc.underlying_biogeochemistry
    redfield, rain = u.plankton.redfield_ratio, u.plankton.carbon_calcite_ratio

    def tracer_N():
        return sum(model.tracers[n].interior.sum().item() for n in ("NO₃", "NH₄", "P", "Z", "sPON", "bPON", "DON"))

    def tracer_C():
        t = model.tracers
        return (sum(t[n].interior.sum().item() for n in ("sPOC", "bPOC", "DOC", "DIC"))
                + (t["P"].interior * (1 + rain) + t["Z"].interior).sum().item() * redfield)

    def kelp_N():
        f = particles.fields
        return (f["A"] * kA * (f["N"] + Ns)).sum().item() / (14 * 0.001)

    def kelp_C():
        f = particles.fields
        return (f["A"] * kA * (f["C"] + Cs)).sum().item() / (12 * 0.001)

    N0, C0, kN0, kC0 = tracer_N(), tracer_C(), kelp_N(), kelp_C()
    model.clock.time = 60 * day  # get to a high growth phase
    for _ in range(10):
        model.time_step(1.0)
    model.finish_particles()
    N1, C1, kN1, kC1 = tracer_N(), tracer_C(), kelp_N(), kelp_C()
    assert kN1 != kN0 and kC1 != kC0  # kelp is being integrated
    eps = lambda x: np.spacing(x)  # noqa: E731
    rtol = max(math.sqrt(eps(N0 + kN0)), math.sqrt(eps(N1 + kN1)))
    assert abs((N0 + kN0) - (N1 + kN1)) <= rtol * max(abs(N0 + kN0), abs(N1 + kN1))
    rtol = max(math.sqrt(eps(C0 + kC0)), math.sqrt(eps(C1 + kC1)))
    assert abs((C0 + kC0) - (C1 + kC1)) <= rtol * max(abs(C0 + kC0), abs(C1 + kC1))
