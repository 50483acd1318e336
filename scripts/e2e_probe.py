#!/usr/bin/env python
"""What limits the host-staged (end-to-end) stage when N ranks of one node run it at once?  Run under torchrun:
every rank builds its y-slab of the pisces_c4 grid (as bench.py does in its default strong mode), then times, between
barriers, a list of variants on ONE set of pinned buffers; rank 0 prints one JSON line per variant with the time of
every rank.  Variants:
  dma:<n> / sm:<n> / sm_d2h:<n>   HostStagedStage with n pipeline slabs and that copy engine
  copies:<n>                      the stage's slab copies alone (no kernels in between), DMA
  h2d:<n> / d2h:<n>               one direction of those copies alone
  wc:<n>                          the stage with its H2D sources in write-combined pinned memory
  flat:<n>                        the same bytes as contiguous chunks (one cudaMemcpyAsync per field and slab) — what a
                                  slab-major host layout would give
usage: torchrun … scripts/e2e_probe.py [variant …]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
from oceanbiome_b200 import _lib
from oceanbiome_b200.host_stage import HostStagedStage, bind_to_gpu_numa_node

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
device = torch.device("cuda", local)
torch.cuda.set_device(device)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
name = os.environ.get("PROBE_WORKLOAD", "pisces_c4")
Ny = bench.workload_table()[name][1]["size"][1]
ny = Ny // world
bind_to_gpu_numa_node(device)
w = bench.Workload(name, device, 1.0, rows=(rank * ny, ny, Ny))
first = HostStagedStage(w.model, nslabs=8)
first.upload_from_device()
bufs = (first.host_tracers, first.host_G)
lib = _lib.load()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)


def gather(ms):
    if world == 1:
        return [ms]
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [round(x.item(), 2) for x in out]


def timed(fn, reps=3):
    for _ in range(2):
        fn()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    barrier()
    return a.elapsed_time(b) / reps


def copies_only(st, h2d=True, d2h=True):
    g = st.grid
    cur = torch.cuda.current_stream(device)

    def fn():
        st.s_in.wait_stream(cur); st.s_out.wait_stream(cur)
        for j0, j1 in st.slabs:
            cg = g.c_grid(j0=j0, j1=j1)
            if h2d:
                _lib.check(lib.obm_copy_slab(C.byref(cg), len(st.names), st._dst_in, st._src_in, st.nplanes, 0, st.s_in.cuda_stream), "h2d")
            if d2h:
                _lib.check(lib.obm_copy_slab(C.byref(cg), len(st.gnames), st._dst_out, st._src_out, st.nplanes, 1, st.s_out.cuda_stream), "d2h")
        cur.wait_stream(st.s_in); cur.wait_stream(st.s_out)
    return fn


def flat(st, n):
    nin, nout = st.h2d_bytes // 8, st.d2h_bytes // 8
    h_in, h_out = torch.empty(nin, dtype=torch.float64).pin_memory(), torch.empty(nout, dtype=torch.float64).pin_memory()
    d_in, d_out = torch.empty(nin, dtype=torch.float64, device=device), torch.zeros(nout, dtype=torch.float64, device=device)
    cin, cout = nin // (n * len(st.names)), nout // (n * len(st.gnames))
    cur = torch.cuda.current_stream(device)

    def fn():
        st.s_in.wait_stream(cur); st.s_out.wait_stream(cur)
        for s in range(n):
            with torch.cuda.stream(st.s_in):
                for f in range(len(st.names)):
                    o = (s * len(st.names) + f) * cin
                    d_in[o:o + cin].copy_(h_in[o:o + cin], non_blocking=True)
            with torch.cuda.stream(st.s_out):
                for f in range(len(st.gnames)):
                    o = (s * len(st.gnames) + f) * cout
                    h_out[o:o + cout].copy_(d_out[o:o + cout], non_blocking=True)
        cur.wait_stream(st.s_in); cur.wait_stream(st.s_out)
    return fn


def write_combined_copies(tensors):
    """The same host arrays in pinned WRITE-COMBINED memory (cudaHostAllocWriteCombined: device reads of it are not snooped
    through the CPU caches) — for the H2D sources only; the CPU must not read such memory back at speed."""
    import ctypes
    import numpy as np
    rt = ctypes.CDLL("libcudart.so.12")
    assert rt.cudaSetDevice(local) == 0
    out, keep = {}, []
    for n, t in tensors.items():
        nbytes = t.numel() * 8
        p = ctypes.c_void_p()
        rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04 | 0x01))  # write-combined | portable
        assert rc == 0, f"cudaHostAlloc → {rc}"
        a = np.ctypeslib.as_array((ctypes.c_double * t.numel()).from_address(p.value)).reshape(tuple(t.shape))
        a[...] = t.numpy()
        out[n] = torch.from_numpy(a)
        keep.append(p)
    return out, keep


specs = sys.argv[1:] or ["dma:8", "dma:1", "dma:2", "dma:4", "dma:16", "copies:8", "copies:2", "h2d:8", "d2h:8", "flat:8", "flat:1",
                         "sm:8", "sm_d2h:8", "dma:8"]
for spec in specs:
    kind, n = spec.split(":")
    n = int(n)
    try:
        if kind == "wc":
            wc, _keep = write_combined_copies(bufs[0])
            st = HostStagedStage(w.model, nslabs=n, host_buffers=(wc, bufs[1]))
            fn = st.step
        elif kind in ("dma", "sm", "sm_h2d", "sm_d2h"):
            st = HostStagedStage(w.model, nslabs=n, copy_engine=kind, host_buffers=bufs)
            fn = st.step
        else:
            st = HostStagedStage(w.model, nslabs=n, host_buffers=bufs)
            fn = {"copies": lambda: copies_only(st), "h2d": lambda: copies_only(st, d2h=False),
                  "d2h": lambda: copies_only(st, h2d=False), "flat": lambda: flat(st, n)}[kind]()
        ms = timed(fn)
        per_rank = gather(ms)
        if rank == 0:
            worst = max(per_rank)
            moved = {"h2d": st.h2d_bytes, "d2h": st.d2h_bytes}.get(kind, max(st.h2d_bytes, st.d2h_bytes))
            print(json.dumps({"variant": spec, "ranks": world, "ms_max": worst, "ms_per_rank": per_rank,
                              "GBs_each_direction_slowest_rank": round(moved / worst / 1e6, 2),
                              "Gcell_s": round(w.cells * world / worst / 1e6, 4)}), flush=True)
        del st, fn
    except Exception as e:  # keep going: the other variants still tell something
        if rank == 0:
            print(json.dumps({"variant": spec, "error": repr(e)}), flush=True)
if world > 1:
    dist.destroy_process_group()
