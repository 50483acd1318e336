#!/usr/bin/env python
"""Measure the many-stream access-pattern ceiling (obm_stream_pattern_gbs) for a few (nread, nrmw) mixes.
usage: stream_pattern.py [Nx Ny Nz]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import oceanbiome_b200 as ob
from oceanbiome_b200 import _lib

N = [int(x) for x in sys.argv[1:4]] or [1024, 128, 128]
dev = torch.device("cuda:0")
grid = ob.RectilinearGrid(size=tuple(N), extent=(1.0, 1.0, 1.0), device=dev)
fields = [ob.CenterField(grid, f"f{n}", fill=1.0) for n in range(64)]
lib = _lib.load()
cg = grid.c_grid()
out = {}
for nread, nrmw, mode in ((1, 1, 1), (8, 10, 0), (8, 10, 1), (22, 0, 0), (0, 22, 0), (37, 24, 0), (37, 24, 1), (37, 0, 0), (0, 24, 0), (0, 24, 1)):
    rd = _lib.pointer_table([f.ptr for f in fields[:nread]] or [fields[0].ptr])
    wr = _lib.pointer_table([f.ptr for f in fields[40:40 + nrmw]] or [fields[40].ptr])
    v = lib.obm_stream_pattern_gbs(C.byref(cg), nread, rd, nrmw, wr, mode, 10, None)
    out[f"read{nread}_{'rmw' if mode == 0 else 'write'}{nrmw}"] = round(v, 1)
print(json.dumps(out))
