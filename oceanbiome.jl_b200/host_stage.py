"""Host-resident stage: the call a user whose tracer fields live in HOST memory makes.

`HostStagedStage(model)` keeps every tracer and every tendency in pinned host arrays (same halo'd parent
layout) and runs one biogeochemical stage — `update_biogeochemical_state!` + all tendencies — for the whole
grid as a 3-stream pipeline over x–y slabs:

        H2D(slab s+1)   ∥   kernels(slab s)   ∥   D2H(slab s−1)

which is legal because every hot kernel is pointwise or column-local.  The PCIe copies, not the
kernels, bound this path; pipelining overlaps the two copy directions and hides the kernels entirely.  With n slabs
the two directions overlap for n − 1 of n + 1 copy slots; measured on B200 (PISCES, 1024-wide grid) n = 8 … 32 all
move 40.4 – 41.5 GB/s per direction (best n = 16, the default) against 49.6 GB/s for two large contiguous copies: the
rows of a slab are one DMA descriptor per (field, k-plane), and that — not fill/drain — is what is left on the table.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .grids import fill_z_halos


def gpu_local_cpus(device) -> Optional[set]:
    """CPUs of the NUMA node the GPU's PCIe root port hangs off (sysfs `local_cpulist`), or None when unknown."""
    try:
        p = torch.cuda.get_device_properties(device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        text = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in text.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus or None
    except (OSError, AttributeError, ValueError):
        return None


def bind_to_gpu_numa_node(device) -> bool:
    """Pin the calling thread to the GPU-local CPUs so that the pinned buffers allocated next are first-touched on the
    GPU's own NUMA node: host↔device copies then do not cross the socket interconnect."""
    import os
    cpus = gpu_local_cpus(device)
    if not cpus:
        return False
    try:
        os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
        return True
    except OSError:
        return False


class HostStagedStage:
    def __init__(self, model, nslabs: int = 16, pin: bool = True, copy_engine: str = "dma", host_buffers=None,
                 return_tracers: bool = False):
        """`return_tracers`: also copy the tracers back after the stage (the state update may have rescaled them —
        `ScaleNegativeTracers` works in place).  Off by default: in the drop-in the tracer arrays live on the device
        (Julia hands over CuArrays) and what a stage produces for its caller is the tendencies."""
        self.model = model
        self.return_tracers = bool(return_tracers)
        self.grid = model.grid
        if copy_engine not in ("sm", "dma", "sm_h2d", "sm_d2h"):
            raise ValueError("copy_engine: 'dma' (cudaMemcpy2DAsync), 'sm' (persistent copy kernel) or 'sm_h2d' / 'sm_d2h' "
                             "(copy kernel in one direction, DMA in the other)")
        if copy_engine != "dma" and not pin:
            raise ValueError("the SM-driven copy needs pinned host buffers")
        self.copy_engine = copy_engine
        g = self.grid
        self.names = list(model.tracers)
        self.gnames = [n for n in self.names if n not in ("T", "S")]
        alloc = (lambda t: torch.empty(t.shape, dtype=torch.float64).pin_memory()) if pin else \
                (lambda t: torch.empty(t.shape, dtype=torch.float64))
        if host_buffers is not None:  # (tracers, tendencies) dicts of host arrays in the parent layout, owned by the caller
            self.host_tracers, self.host_G = host_buffers
        else:
            self.host_tracers = {n: alloc(model.tracers[n].data) for n in self.names}
            self.host_G = {n: alloc(model.Gn[n].data) for n in self.gnames}
        nslabs = max(1, min(nslabs, g.Ny))
        edges = [round(s * g.Ny / nslabs) for s in range(nslabs + 1)]
        self.slabs = [(edges[s], edges[s + 1]) for s in range(nslabs) if edges[s + 1] > edges[s]]
        dev = g.device
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        # only the Nz interior k-planes travel: no hot kernel reads a tracer halo (they are pointwise / column-local
        # over interior cells) and tendencies are written to interior cells only
        self.nplanes = g.Nz
        skip = g.Hz * (g.Nx + 2 * g.Hx) * (g.Ny + 2 * g.Hy) * 8  # bytes of the bottom halo planes
        plane_bytes = (g.Nx + 2 * g.Hx) * 8 * self.nplanes
        self.h2d_bytes = sum((j1 - j0) for j0, j1 in self.slabs) * plane_bytes * len(self.names)
        self.d2h_bytes = sum((j1 - j0) for j0, j1 in self.slabs) * plane_bytes * (
            len(self.gnames) + (len(self.names) if self.return_tracers else 0))
        sed = getattr(model.biogeochemistry, "sediment", None)
        self._halo_names = list(sed.biogeochemistry.sinking_fluxes()) if sed is not None else []
        self._src_in = _lib.pointer_table([self.host_tracers[n].data_ptr() + skip for n in self.names])
        self._dst_in = _lib.pointer_table([model.tracers[n].ptr + skip for n in self.names])
        self._src_out = _lib.pointer_table([model.Gn[n].ptr + skip for n in self.gnames])
        self._dst_out = _lib.pointer_table([self.host_G[n].data_ptr() + skip for n in self.gnames])

    def upload_from_device(self):
        """Initialise the host copies from the model's current device state."""
        for n in self.names:
            self.host_tracers[n].copy_(self.model.tracers[n].data)

    def step(self, kernels: bool = True, h2d: bool = True, d2h: bool = True):
        """One stage, host → host.  Returns after all work is enqueued; `synchronize()` to wait.
        `kernels=False` / `h2d=False` / `d2h=False` leave that part of the pipeline out: the same copies of the same
        buffers without the work between them (or one direction of them) are the measured ceiling of this path on the
        box at hand (bench.py's e2e roofline) — results are only meaningful with all three on."""
        lib = _lib.load()
        copy_in = lib.obm_copy_slab_sm if self.copy_engine in ("sm", "sm_h2d") else lib.obm_copy_slab
        copy_out = lib.obm_copy_slab_sm if self.copy_engine in ("sm", "sm_d2h") else lib.obm_copy_slab
        m, g = self.model, self.grid
        bgc = m.biogeochemistry
        cur = torch.cuda.current_stream(g.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        for j0, j1 in self.slabs:
            cg = g.c_grid(j0=j0, j1=j1)
            if h2d:
                rc = copy_in(C.byref(cg), len(self.names), self._dst_in, self._src_in, self.nplanes, 0,
                             self.s_in.cuda_stream)
                _lib.check(rc, "obm_copy_slab(H2D)")
            ready = self.s_in.record_event()
            self.s_run.wait_event(ready)
            if kernels:
                with torch.cuda.stream(self.s_run), g.restrict(j0, j1):
                    # only interior planes travel; the one halo a hook reads — the plane below the bottom cells of the
                    # sediment's sinking tracers — is rebuilt on the device for this slab's rows (zero gradient, as
                    # Oceananigans' fill_halo_regions! leaves it)
                    for n in self._halo_names:
                        fill_z_halos(m.tracers[n], j0, j1)
                    bgc.update_biogeochemical_state(m)
                    bgc.underlying_biogeochemistry.compute_tendencies(g, m.tracers, bgc.biogeochemical_auxiliary_fields(),
                                                                     m.Gn, accumulate=False, time=m.clock.time)
                    if bgc.sediment is not None:
                        bgc.sediment.update_tendencies(bgc, m, None)
            done = self.s_run.record_event()
            self.s_out.wait_event(done)
            if not d2h:
                continue
            rc = copy_out(C.byref(cg), len(self.gnames), self._dst_out, self._src_out, self.nplanes, 1,
                          self.s_out.cuda_stream)
            _lib.check(rc, "obm_copy_slab(D2H)")
            if self.return_tracers:  # src / dst swapped: device tracers → host tracers
                rc = copy_out(C.byref(cg), len(self.names), self._src_in, self._dst_in, self.nplanes, 1,
                              self.s_out.cuda_stream)
                _lib.check(rc, "obm_copy_slab(D2H, tracers)")
        for s in (self.s_in, self.s_run, self.s_out):
            cur.wait_stream(s)

    def synchronize(self):
        torch.cuda.current_stream(self.grid.device).synchronize()
