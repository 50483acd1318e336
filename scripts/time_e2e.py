#!/usr/bin/env python
"""Time the host-staged (end-to-end) stage for several pipeline configurations on one set of pinned buffers.
usage: time_e2e.py <workload> <scale> <engine:nslabs> [<engine:nslabs> ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oceanbiome_b200.host_stage import HostStagedStage

name, scale = sys.argv[1], float(sys.argv[2])
w = bench.Workload(name, torch.device("cuda:0"), scale)
first = HostStagedStage(w.model, nslabs=8)
first.upload_from_device()
bufs = (first.host_tracers, first.host_G)
for spec in sys.argv[3:]:
    if spec == "numa":  # re-allocate the pinned buffers on the GPU's NUMA node
        from oceanbiome_b200.host_stage import bind_to_gpu_numa_node, gpu_local_cpus
        print(json.dumps({"gpu_local_cpus": len(gpu_local_cpus(0) or ()), "bound": bind_to_gpu_numa_node(0),
                          "nodes": sorted(os.listdir("/sys/devices/system/node"))[:8]}), flush=True)
        del first, bufs
        first = HostStagedStage(w.model, nslabs=8)
        first.upload_from_device()
        bufs = (first.host_tracers, first.host_G)
        continue
    engine, nslabs = spec.split(":")
    st = HostStagedStage(w.model, nslabs=int(nslabs), copy_engine=engine, host_buffers=bufs)
    for _ in range(2):
        st.step()
    st.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        st.step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(json.dumps({"engine": engine, "nslabs": int(nslabs), "ms": ms, "Gcell_s": w.cells / ms / 1e6,
                      "GBs_each_direction": max(st.h2d_bytes, st.d2h_bytes) / ms / 1e6}), flush=True)
