#!/usr/bin/env python
"""Generate tests/golden/npd_tendencies.json: tendencies of the Nutrients–Plankton–Detritus family at seeded states,
evaluated by the independent Python transliteration of the reference (oracle/pyref_npd.py).  The C oracle must
reproduce them (tests/test_oracle_npd.py::test_c_oracle_matches_independent_restatement).
usage: python scripts/make_npd_golden.py > tests/golden/npd_tendencies.json"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref_npd as ref  # noqa: E402

CASES = {
    # name: (model factory, tracer order = required_biogeochemical_tracers, carbonate?, oxygen?)
    "lobster": (lambda: ref.lobster(), ["NO₃", "NH₄", "P", "Z", "sPOM", "bPOM", "DOM"], False, False),
    "lobster_carbonate_oxygen": (lambda: ref.lobster(), ["NO₃", "NH₄", "P", "Z", "sPOM", "bPOM", "DOM", "DIC", "Alk", "O₂"], True, True),
    "lobster_iron_variable_redfield_carbonate_oxygen": (
        lambda: ref.lobster("VariableRedfieldDetritus", iron=True),
        ["NO₃", "NH₄", "Fe", "P", "Z", "sPOC", "bPOC", "DOC", "sPON", "bPON", "DON", "DIC", "Alk", "O₂"], True, True),
    "npzd": (lambda: ref.npzd(), ["N", "P", "Z", "T", "D"], False, False),
    "npzd_carbonate_oxygen": (lambda: ref.npzd(), ["N", "P", "Z", "T", "D", "DIC", "Alk", "O₂"], True, True),
}


def state(rng):
    v = {"P": rng.uniform(0.01, 1.5), "Z": rng.uniform(0.01, 1.0), "NO₃": rng.uniform(0.0, 12.0), "NH₄": rng.uniform(0.0, 1.5),
         "Fe": rng.uniform(0.0, 1e-3), "N": rng.uniform(0.0, 12.0), "DIC": rng.uniform(1900, 2300), "Alk": rng.uniform(2100, 2500),
         "O₂": rng.uniform(150, 350), "sPOM": rng.uniform(0, 1), "bPOM": rng.uniform(0, 1), "DOM": rng.uniform(0, 1),
         "D": rng.uniform(0, 1), "T": rng.uniform(2, 25), "PAR": rng.uniform(0.0, 150.0)}
    for n in ("sPON", "bPON", "DON"):
        v[n] = rng.uniform(0, 1)
    for n, m in (("sPOC", "sPON"), ("bPOC", "bPON"), ("DOC", "DON")):
        v[n] = v[m] * rng.uniform(5.0, 9.0)  # off-Redfield on purpose
    return v


out = {"generator": "scripts/make_npd_golden.py (oracle/pyref_npd.py)", "cases": {}}
rng = random.Random(20261017)
for name, (factory, tracers, _, _) in CASES.items():
    model = factory()
    rows = []
    for _ in range(6):
        f = state(rng)
        rows.append({"state": {n: f[n] for n in tracers + ["PAR"]},
                     "tendencies": {n: (0.0 if n == "T" else model(n, f)) for n in tracers}})
    # a state with zeros where the eps(0.0) guards act
    f = state(rng)
    f.update({"P": 0.0, "sPOM": 0.0, "sPON": 0.0, "D": 0.0, "NO₃": 0.0, "NH₄": 0.0, "N": 0.0})
    rows.append({"state": {n: f[n] for n in tracers + ["PAR"]},
                 "tendencies": {n: (0.0 if n == "T" else model(n, f)) for n in tracers}})
    out["cases"][name] = {"tracers": tracers, "rows": rows}
json.dump(out, sys.stdout, ensure_ascii=False, indent=1)
