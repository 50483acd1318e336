#!/usr/bin/env python
"""Ω kernel: cold start vs warm start (static state) vs warm start under a realistic per-stage drift of DIC/Alk."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
w = bench.Workload("pisces_c4", torch.device("cuda:0"), 0.125)
m = w.model
u = m.biogeochemistry.underlying_biogeochemistry
t = m.tracers
def timeit(fn, n=10, pre=None):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if pre: pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n
cc = u.carbon_chemistry
import oceanbiome_b200 as ob
state = ob.CenterField(m.grid, "lnH")
call = lambda st: cc.calcite_saturation(m.grid, t["T"], t["S"], t["DIC"], t["Alk"], t["Si"], u.calcite_saturation, state=st)
res = {"cold_ms": timeit(lambda: call(None))}
om_cold = u.calcite_saturation.data.clone()
res["warm_static_ms"] = timeit(lambda: call(state))
res["warm_vs_cold_max_rel"] = ((u.calcite_saturation.data - om_cold).abs() / om_cold.abs().clamp_min(1e-300)).max().item()
for drift in (1e-5, 1e-4, 1e-3):
    def pre(d=drift):
        t["DIC"].data.mul_(1 + d); t["Alk"].data.mul_(1 - 0.3 * d)
    res[f"warm_drift_{drift:g}_ms"] = timeit(lambda: call(state), pre=pre)
print(json.dumps(res))
