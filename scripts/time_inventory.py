#!/usr/bin/env python
"""Time the tracer-inventory kernel (obm_inventory: the local part of the path's one collective) on the PISCES grid.
usage: time_inventory.py [scale]   (library selected by $OBM_B200_LIB)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oceanbiome_b200.distributed import InventoryDiagnostic

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
w = bench.Workload("pisces_c4", torch.device("cuda:0"), scale)
diag = InventoryDiagnostic(w.grid, w.model.tracers, w.groups)
for _ in range(3):
    diag.local()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    diag.local()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(json.dumps({"lib": os.environ.get("OBM_B200_LIB", "default"), "cells": w.cells, "ms": round(ms, 4),
                  "GBs": round(8 * len(diag.names) * w.cells / ms / 1e6, 1), "totals": [float(x) for x in diag.local().cpu()]}))
