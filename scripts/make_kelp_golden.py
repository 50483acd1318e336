#!/usr/bin/env python
"""Generate tests/golden/kelp_rates.json: the eleven sugar-kelp rates (A, N, C and the eight coupled tracers) at seeded
states, evaluated by the independent Python transliteration of the reference (oracle/pyref_kelp.py).  The C oracle must
reproduce them (tests/test_oracle_kelp.py::test_c_oracle_matches_independent_restatement).
usage: python scripts/make_kelp_golden.py > tests/golden/kelp_rates.json"""
import json
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref_kelp as ref  # noqa: E402

day = 86400.0
rng = random.Random(20261017)
# the reference test's own state (test/test_sugar_kelp.jl: N far above N_max) first, then seeded states that take both
# sides of every min / max / comparison: cold and warm of the optimal range, reserves near their minima, fast and no flow
states = [dict(t=60 * day, A=2.0, N=1.0, C=1.0, u=0.0, v=0.0, w=0.0, T=10.0, NO3=10.0, NH4=1.0, PAR=50.0)]
for _ in range(11):
    states.append(dict(t=rng.uniform(0, 800) * day, A=rng.choice([0.3, 3.0, 30.0]) * rng.uniform(0.5, 1.5),
                       N=rng.uniform(0.0127, 0.03), C=rng.uniform(0.0101, 0.7), u=rng.choice([0.0, 0.02, 0.3]) * rng.random(),
                       v=rng.uniform(-0.1, 0.1), w=rng.uniform(-0.01, 0.01), T=rng.uniform(-1.5, 22.0),
                       NO3=rng.uniform(0.0, 15.0), NH4=rng.uniform(0.0, 4.0), PAR=rng.uniform(0.0, 250.0)))
out = {"generator": "scripts/make_kelp_golden.py (oracle/pyref_kelp.py)", "names": list(ref.NAMES), "cases": []}
for CN, kw in (("Inf", {}), ("12", {"exudation_redfield_ratio": 12.0}),
               ("12, warm-adapted", {"exudation_redfield_ratio": 12.0, "adapted_latitude": 45.0, "erosion_exponent": 0.3})):
    kelp = ref.SugarKelp(**kw)
    for s in states:
        rates = [kelp(n, s["t"], s["A"], s["N"], s["C"], s["u"], s["v"], s["w"], s["T"], s["NO3"], s["NH4"], s["PAR"])
                 for n in ref.NAMES]
        assert all(math.isfinite(r) for r in rates)
        out["cases"].append({"parameters": kw, "state": s, "rates": rates})
json.dump(out, sys.stdout, indent=1, ensure_ascii=False)
