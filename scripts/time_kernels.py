#!/usr/bin/env python
"""Time individual hot-path kernels with CUDA events (library selected by $OBM_B200_LIB).
usage: time_kernels.py <workload> [scale]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "pisces_c4"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.125
dev = torch.device("cuda:0")
w = bench.Workload(name, dev, scale)
m = w.model
bgc = m.biogeochemistry
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
res = {"lib": os.environ.get("OBM_B200_LIB", "default"), "cells": w.cells}
from oceanbiome_b200.biogeochemistry import _update_modifiers
res["scale_negative_ms"] = timeit(lambda: _update_modifiers(m, bgc.modifiers, None))
if hasattr(bgc.underlying_biogeochemistry, "calcite_saturation_arguments"):
    u0 = bgc.underlying_biogeochemistry
    res["scale_negative_calcite_fused_ms"] = timeit(lambda: _update_modifiers(m, bgc.modifiers, None, u0.calcite_saturation_arguments(m)))
res["light_ms"] = timeit(lambda: bgc.light_attenuation.update_biogeochemical_state(m))
if hasattr(bgc.underlying_biogeochemistry, "column_light_state"):
    res["light_with_column_state_ms"] = timeit(lambda: bgc.light_attenuation.update_biogeochemical_state(m, column_state=bgc.underlying_biogeochemistry.column_light_state(m)))
res["underlying_state_ms"] = timeit(lambda: bgc.underlying_biogeochemistry.update_biogeochemical_state(m))
res["tendencies_ms"] = timeit(lambda: bgc.update_tendencies(m))
res["tendency_Gcell_s"] = w.cells / res["tendencies_ms"] / 1e6
u = bgc.underlying_biogeochemistry
res["tendencies_overwrite_ms"] = timeit(lambda: u.compute_tendencies(m.grid, m.tracers, bgc.biogeochemical_auxiliary_fields(), m.Gn,
                                                                      accumulate=False, time=m.clock.time))
if len(sys.argv) > 3 and sys.argv[3] == "carbon":
    wc = bench.Workload("carbon_c5", dev, 0.25)
    res["carbon_sweep_ms_25M"] = timeit(lambda: wc.step())
print(json.dumps(res))
