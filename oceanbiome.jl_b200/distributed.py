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


def local_inventory(grid: RectilinearGrid, tracers: dict, groups: Sequence, stream: Optional[int] = None) -> torch.Tensor:
    """Per-GPU fused reduction out[g] = Σ_cells (Σ_f sf·c_f)·V (obm_inventory).  `groups` is a list of
    (tracer_names, scalefactors); returns a device tensor of len(groups) doubles."""
    names, garr = inventory_groups(groups)
    fields = [tracers[n] for n in names]
    require_cuda(*fields)
    lib = _lib.load()
    dev = grid.device
    out = torch.zeros(len(groups), dtype=torch.float64, device=dev)
    ws = torch.empty(lib.obm_inventory_workspace_bytes(len(groups)) // 8, dtype=torch.float64, device=dev)
    vol = torch.zeros(grid.parent_shape, dtype=torch.float64, device=dev)
    grid.interior(vol)[...] = grid.cell_volume()
    cg = grid.c_grid()
    s = stream if stream is not None else current_stream_ptr(dev)
    rc = lib.obm_inventory(C.byref(cg), len(names), _lib.pointer_table([f.ptr for f in fields]), len(groups), garr,
                           vol.data_ptr(), 0.0, out.data_ptr(), ws.data_ptr(), s)
    _lib.check(rc, "obm_inventory")
    return out


def allreduce_sum(values: torch.Tensor) -> torch.Tensor:
    """The one collective of the path: sum of the per-slab inventories (NCCL on GPUs, gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(values, op=dist.ReduceOp.SUM)
    return values


def tracer_inventory(grid: RectilinearGrid, tracers: dict, groups: Sequence, stream: Optional[int] = None) -> torch.Tensor:
    """Global inventory: local fused reduction + ONE all-reduce (sum) over the slabs."""
    return allreduce_sum(local_inventory(grid, tracers, groups, stream))
