#!/usr/bin/env python
"""Straight-line vs looped instruction stream (obm_fetch_ceiling_ms): same 4 096 DFMA per thread."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oceanbiome_b200 import _lib
lib = _lib.load()
out = {}
for blocks in (148 * 4 * 8, 148 * 4 * 64):
    s = torch.empty(blocks * 128, dtype=torch.float64, device="cuda")
    for straight in (0, 1):
        out[f"blocks{blocks}_{'straight' if straight else 'loop'}_ms"] = round(min(lib.obm_fetch_ceiling_ms(s.data_ptr(), blocks, straight, None) for _ in range(3)), 4)
print(json.dumps(out))
