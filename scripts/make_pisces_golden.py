#!/usr/bin/env python
"""Generate tests/golden/pisces_tendencies.json: the 24 PISCES tendencies at seeded states, evaluated by the independent
Python transliteration of the reference (oracle/pyref_pisces.py).  The states are drawn so that both sides of every
branch occur: above / below the mixed layer and the euphotic depth, Ω below and above 1, a southern latitude (enhanced
silicate uptake), oxic / anoxic water, sinking on (flux feeding) and off, zero biomass (the eps(0.0) guards).
usage: python scripts/make_pisces_golden.py > tests/golden/pisces_tendencies.json"""
import json
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref_pisces as ref  # noqa: E402

INITIAL = {"P": 0.5, "PChl": 0.02, "PFe": 0.005, "D": 0.1, "DChl": 0.004, "DFe": 0.001, "DSi": 0.01, "Z": 0.1, "M": 0.7,
           "DOC": 2.1, "POC": 7.8, "SFe": 0.206, "GOC": 38.0, "BFe": 1.1, "PSi": 0.1, "CaCO₃": 0.3, "NO₃": 2.3,
           "NH₄": 0.9, "PO₄": 0.6, "Fe": 0.13, "Si": 8.5, "DIC": 2205.0, "Alk": 2566.0, "O₂": 317.0}
rng = random.Random(20261018)
rows = []
for case in range(14):
    f = {n: v * math.exp(rng.uniform(-1.0, 1.0)) for n, v in INITIAL.items()}
    f["T"] = rng.uniform(-1.0, 29.0)
    f["S"] = rng.uniform(33.0, 37.0)
    f["PAR₁"], f["PAR₂"], f["PAR₃"] = (rng.uniform(0.0, 60.0) for _ in range(3))
    f["PAR"] = f["PAR₁"] + f["PAR₂"] + f["PAR₃"]
    f["Ω"] = rng.choice([0.35, 0.8, 1.7, 4.2])
    f["zₘₓₗ"] = -rng.uniform(10.0, 150.0)
    f["zₑᵤ"] = -rng.uniform(20.0, 120.0)
    f["z"] = -rng.uniform(1.0, 300.0)
    f["κ"] = 10 ** rng.uniform(-4, -1)
    f["mixed_layer_PAR"] = rng.uniform(0.0, 80.0)
    f["wPOC"], f["wGOC"] = -2.0 / 86400, -rng.uniform(30.0, 200.0) / 86400
    f["Si_clim"] = 7.5
    f["t"] = rng.choice([1.6, 0.37 * 365 * 86400.0, 2.0e7, 86400.0 * 200.3])
    lat = rng.choice([45.0, -52.5, 12.0])
    if case == 3:
        f["O₂"] = 0.8          # anoxic: ΔO₂ = 1
    if case == 4:
        f["O₂"] = 4.0          # partly anoxic
    if case == 5:
        f["wPOC"] = f["wGOC"] = 0.0
    if case == 6:
        f.update({"P": 0.0, "PChl": 0.0, "PFe": 0.0, "POC": 0.0, "SFe": 0.0})   # eps(0.0) guards
    if case == 7:
        f.update({"D": 0.0, "DChl": 0.0, "DFe": 0.0, "DSi": 0.0, "GOC": 0.0, "BFe": 0.0, "DOC": 0.0})
    if case == 8:
        f.update({"PAR₁": 0.0, "PAR₂": 0.0, "PAR₃": 0.0, "PAR": 0.0, "mixed_layer_PAR": 0.0})  # dark
    model = ref.PISCES(latitude=lat)
    rows.append({"latitude": lat, "state": f, "tendencies": {n: model(n, f) for n in ref.TRACERS},
                 "day_length_growth": ref.cbm_day_length(lat, f["t"]), "day_length_chlorophyll": ref.cbm_day_length(f["t"], lat)})
json.dump({"generator": "scripts/make_pisces_golden.py (oracle/pyref_pisces.py)", "rows": rows}, sys.stdout, ensure_ascii=False, indent=1)
