"""Reference-held goldens, when somebody with Julia has produced them: `bench_ref/dump_goldens.jl` evaluates the REAL
OceanBioME.jl per-tracer callables (`bgc(i, j, k, grid, Val(name), clock, fields, auxiliary_fields)`, PISCES.jl:122-123,
NutrientsPlanktonDetritus.jl:88) at the states of tests/golden/{pisces,npd}_tendencies.json and writes
tests/golden/reference_{pisces,npd}_tendencies.json in the same schema.  The build image has no Julia, so those files are
absent here and the comparisons are skipped (stated in DESIGN.md §4: absolute tendency values rest on two independent
readings of the source until then); with the files present the oracle is pinned on reference OUTPUT:
|oracle − reference| ≤ 1e-13 · max(|reference|, Σ|terms| of that tendency)."""
import json
import math
import os

import numpy as np
import pytest

import oceanbiome_b200 as ob
from oceanbiome_b200 import pisces

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REF_PISCES = os.path.join(GOLDEN, "reference_pisces_tendencies.json")
REF_NPD = os.path.join(GOLDEN, "reference_npd_tendencies.json")


def test_dump_script_reads_the_states_this_repository_holds():
    """The Julia script and this test agree on file names and schema (checked without Julia: the script names the two
    input files, the two output files, and every key of a row)."""
    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "bench_ref", "dump_goldens.jl"), encoding="utf-8").read()
    for name in ("pisces_tendencies.json", "npd_tendencies.json", "reference_pisces_tendencies.json", "reference_npd_tendencies.json"):
        assert name in src
    row = json.load(open(os.path.join(GOLDEN, "pisces_tendencies.json"), encoding="utf-8"))["rows"][0]
    for key in list(row) + ["Si_clim", "zₘₓₗ", "zₑᵤ", "mixed_layer_PAR", "wPOC", "wGOC", "Ω", "κ", "PAR₁", "PAR₂", "PAR₃"]:
        assert f'"{key}"' in src, key


@pytest.mark.skipif(not os.path.exists(REF_PISCES), reason="no reference-produced PISCES goldens (bench_ref/dump_goldens.jl needs Julia)")
def test_pisces_oracle_against_reference_output(oracle):
    ref = json.load(open(REF_PISCES, encoding="utf-8"))
    mine = json.load(open(os.path.join(GOLDEN, "pisces_tendencies.json"), encoding="utf-8"))
    assert len(ref["rows"]) == len(mine["rows"])
    g = ob.RectilinearGrid(size=(1,), z=(-10, 0), topology=("Flat", "Flat", "Bounded"), device="cpu")
    for r, m in zip(ref["rows"], mine["rows"]):
        f = m["state"]
        assert r["state"] == f, "the reference file was produced from other states"
        u = ob.PISCES(g, latitude=pisces.PrescribedLatitude(m["latitude"])).underlying_biogeochemistry
        p = u.c_params(f["t"])
        assert math.isclose(p.day_length_growth, r["day_length_growth"], rel_tol=1e-13)
        assert math.isclose(p.day_length_chlorophyll, r["day_length_chlorophyll"], rel_tol=1e-13)
        got, S = oracle.pisces_point_terms(p, [f.get(n, 0.0) for n in pisces.TRACERS], f["PAR₁"], f["PAR₂"], f["PAR₃"], f["PAR"],
                                           f["Ω"], f["wPOC"], f["wGOC"], f["zₘₓₗ"], f["zₑᵤ"], f["κ"], f["mixed_layer_PAR"], f["z"])
        for q, n in enumerate(pisces.TRACERS[:24]):
            want = r["tendencies"][n]
            assert abs(got[q] - want) <= 1e-13 * max(abs(want), S[q]), (n, got[q], want)


@pytest.mark.skipif(not os.path.exists(REF_NPD), reason="no reference-produced NPD goldens (bench_ref/dump_goldens.jl needs Julia)")
def test_npd_oracle_against_reference_output(oracle):
    from test_oracle_npd import GOLDEN_MODELS
    ref = json.load(open(REF_NPD, encoding="utf-8"))["cases"]
    mine = json.load(open(os.path.join(GOLDEN, "npd_tendencies.json"), encoding="utf-8"))["cases"]
    one = oracle.Grid.like(ob.RectilinearGrid(size=(1, 1, 1), extent=(1, 1, 2), device="cpu"))
    for case, c in mine.items():
        bgc = GOLDEN_MODELS[case]()
        names = list(bgc.required_biogeochemical_tracers())
        assert names == ref[case]["tracers"]
        for r, m in zip(ref[case]["rows"], c["rows"]):
            assert r["state"] == m["state"], "the reference file was produced from other states"
            tr = []
            for n in names:
                a = np.zeros(one.parent_shape)
                one.interior(a)[...] = m["state"][n]
                tr.append(a)
            par = np.full(one.parent_shape, m["state"]["PAR"])
            G = oracle.npd_tendencies(one, bgc.c_params(), tr, par)
            S = oracle.npd_tendency_scales(one, bgc.c_params(), tr, par)
            for n, g, s in zip(names, G, S):
                got, sc, want = float(one.interior(g)[0, 0, 0]), float(one.interior(s)[0, 0, 0]), r["tendencies"][n]
                assert abs(got - want) <= 1e-13 * max(abs(want), sc), (case, n, got, want)
