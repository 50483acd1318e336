"""Multi-GPU plumbing: one process per GPU, each owning an x–y slab of the grid
(`RectilinearGrid.slab(rank, world)`).  Every hot kernel is pointwise or column-local, so the data
path needs NO collective; the only exchange is the all-reduce of the per-GPU tracer inventories
(conservation diagnostics) — a handful of doubles over NCCL / NVLink.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from .grids import RectilinearGrid, current_stream_ptr, require_cuda


def init_distributed(backend: Optional[str] = None):
    """Rendezvous from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world, device).  Single-process when WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    device = torch.device(f"cuda:{local}") if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {"device_id": device} if cuda else {}
        dist.init_process_group(backend or ("nccl" if cuda else "gloo"), rank=rank, world_size=world, **kwargs)
    return rank, world, device


def slab_ranges(Ny: int, world: int):
    """[j0, j1) of every rank's slab; Ny must divide evenly (equal work per GPU)."""
    if Ny % world:
        raise ValueError(f"Ny = {Ny} is not divisible by {world} ranks")
    n = Ny // world
    return [(r * n, (r + 1) * n) for r in range(world)]


def slab_ranges_by_rate(Ny: int, rates: Sequence[float], minimum: int = 1):
    """[j0, j1) of every rank's slab with the rows shared out in proportion to `rates` (largest-remainder rounding, at
    least `minimum` rows each).  For the HOST-buffer path on a node whose GPUs do not see the same host bandwidth
    (measured on the 8-GPU box: the four GPUs behind one host bridge move 1.43× what the other four do,
    profiles/r04_e2e_probe_n8.jsonl) a stage ends when the slowest link has finished, so equal slabs leave the fast links
    idle; the device-resident path keeps `slab_ranges` (equal work per GPU)."""
    world = len(rates)
    if world < 1 or Ny < world * minimum or any(not (r > 0) for r in rates):
        raise ValueError(f"cannot share {Ny} rows over rates {list(rates)} with at least {minimum} each")
    total = float(sum(rates))
    exact = [Ny * r / total for r in rates]
    rows = [int(e) for e in exact]
    for q in sorted(range(world), key=lambda q: exact[q] - rows[q], reverse=True)[:Ny - sum(rows)]:
        rows[q] += 1
    while min(rows) < minimum:  # a very slow rank still gets `minimum` rows, taken from the largest slab
        rows[rows.index(max(rows))] -= 1
        rows[rows.index(min(rows))] += 1
    out, j = [], 0
    for q in range(world):
        out.append((j, j + rows[q]))
        j += rows[q]
    assert j == Ny
    return out


def inventory_groups(groups: Sequence):
    """→ (distinct tracer names, ctypes obm_scale_group array) for a list of (names, scalefactors)."""
    names = []
    for tn, _ in groups:
        for t in tn:
            if t not in names:
                names.append(t)
    garr = (_lib.obm_scale_group * len(groups))()
    for q, (tn, sf) in enumerate(groups):
        garr[q].n = len(tn)
        for m, (t, f) in enumerate(zip(tn, sf)):
            garr[q].index[m] = names.index(t)
            garr[q].scalefactor[m] = float(f)
    return names, garr


class InventoryDiagnostic:
    """The conservation diagnostic as a reusable object: buffers (output, reduction workspace, the cell-volume array a
    stretched grid needs) are allocated once, so calling it every stage costs one `obm_inventory` (one fused pass over
    the cells for all groups, HBM-bound) plus ONE all-reduce of `len(groups)` doubles — the only collective of the path.
    `groups` is a list of (tracer_names, scalefactors), e.g. the five PISCES element budgets of
    `conserved_tracers` (PISCES/coupling_utils.jl:10-33) or the N / C budgets the reference's tests sum by hand
    (test/test_NutrientsPlanktonDetritus.jl:8-21)."""

    def __init__(self, grid: RectilinearGrid, tracers: dict, groups: Sequence):
        import numpy as np
        self.grid, self.groups = grid, list(groups)
        self.names, self.garr = inventory_groups(groups)
        self.fields = [tracers[n] for n in self.names]
        require_cuda(*self.fields)
        lib = _lib.load()
        dev = grid.device
        self.out = torch.zeros(len(self.groups), dtype=torch.float64, device=dev)
        self.workspace = torch.empty(lib.obm_inventory_workspace_bytes(len(self.groups)) // 8, dtype=torch.float64, device=dev)
        dz = grid.dz
        if np.allclose(dz, dz[0], rtol=1e-14, atol=0.0):
            self.volume, self.uniform_volume = None, float(dz[0] * grid.dx * grid.dy)
        else:  # stretched z: per-cell volumes, built once
            self.volume = torch.zeros(grid.parent_shape, dtype=torch.float64, device=dev)
            grid.interior(self.volume)[...] = grid.cell_volume()
            self.uniform_volume = 0.0
        self.table = _lib.pointer_table([f.ptr for f in self.fields])

    def local(self, stream: Optional[int] = None) -> torch.Tensor:
        """This GPU's slab: out[g] = Σ_cells (Σ_f sf·c_f)·V (device tensor, overwritten by every call)."""
        cg = self.grid.c_grid()
        s = stream if stream is not None else current_stream_ptr(self.grid.device)
        rc = _lib.load().obm_inventory(C.byref(cg), len(self.names), self.table, len(self.groups), self.garr,
                                       self.volume.data_ptr() if self.volume is not None else None, self.uniform_volume,
                                       self.out.data_ptr(), self.workspace.data_ptr(), s)
        _lib.check(rc, "obm_inventory")
        return self.out

    def __call__(self, stream: Optional[int] = None) -> torch.Tensor:
        """Global inventory: the local reduction, then the all-reduce (sum) over the slabs."""
        return allreduce_sum(self.local(stream))


def local_inventory(grid: RectilinearGrid, tracers: dict, groups: Sequence, stream: Optional[int] = None) -> torch.Tensor:
    """Per-GPU fused reduction out[g] = Σ_cells (Σ_f sf·c_f)·V (obm_inventory).  `groups` is a list of
    (tracer_names, scalefactors); returns a device tensor of len(groups) doubles.  One-off form of
    `InventoryDiagnostic` (which keeps its buffers between calls)."""
    return InventoryDiagnostic(grid, tracers, groups).local(stream).clone()


def allreduce_sum(values: torch.Tensor) -> torch.Tensor:
    """The one collective of the path: sum of the per-slab inventories (NCCL on GPUs, gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(values, op=dist.ReduceOp.SUM)
    return values


def tracer_inventory(grid: RectilinearGrid, tracers: dict, groups: Sequence, stream: Optional[int] = None) -> torch.Tensor:
    """Global inventory: local fused reduction + ONE all-reduce (sum) over the slabs."""
    return allreduce_sum(local_inventory(grid, tracers, groups, stream))
