"""GPU, N = 2 ranks over NCCL (skipped on a box with one GPU): the x–y slab decomposition of the PISCES stage and the
path's ONLY collective — the all-reduce of the per-slab tracer inventories.

  * every rank owns Ny / 2 rows of ONE global grid and fills them with ITS rows of the global synthetic field;
  * obm_inventory on each slab + `all_reduce(SUM)` over NCCL reproduces the oracle's serial sum over the global grid to
    1e-12 · Σ|terms| (five PISCES element budgets with their scale factors), identically on both ranks and run to run;
  * the stage itself needs no exchange: the tendencies a rank computes for its slab are bit for bit the rows the same
    kernels produce on the undivided grid (every kernel is pointwise or column-local)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZE, EXTENT = (48, 16, 12), (4800.0, 1600.0, 300.0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(ob, grid, rows):
    from oceanbiome_b200 import pisces, synthetic
    bgc = ob.PISCES(grid, scale_negatives=True, surface_photosynthetically_active_radiation=100.0)
    bgc.underlying_biogeochemistry.warm_start_carbonate_solve = False
    model = ob.BiogeochemicalModel(grid, bgc)
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *pisces.synthetic_range(n), rows=rows)
    u = bgc.underlying_biogeochemistry
    u.mixed_layer_depth.data.fill_(-80.0)  # column fields identical on every rank: the slab result must not depend on them
    u.mean_mixed_layer_vertical_diffusivity.data.fill_(1e-3)
    u.euphotic_depth.data.fill_(-60.0)
    u.sinking_velocities["GOC"] = pisces.DepthDependantSinkingSpeed().face_field(grid, u.mixed_layer_depth, u.euphotic_depth)
    return bgc, model


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import oceanbiome_b200 as ob
    from oceanbiome_b200 import distributed, pisces
    ob.load_library()
    r, w, dev = distributed.init_distributed()
    assert (r, w) == (rank, world) and dev.type == "cuda"
    full = ob.RectilinearGrid(size=SIZE, extent=EXTENT, device=dev)
    slab = full.slab(rank, world)
    j0, j1 = distributed.slab_ranges(full.Ny, world)[rank]
    bgc, model = _build(ob, slab, rows=(j0, full.Ny))
    groups = [(m.tracers, m.scalefactors) for m in bgc.modifiers]
    diag = distributed.InventoryDiagnostic(slab, model.tracers, groups)
    local = diag.local().clone()
    total = diag().clone()          # obm_inventory + all_reduce(SUM) over NCCL
    again = diag().clone()
    model.update_state()
    model.compute_tendencies()
    torch.cuda.synchronize()
    G = {n: model.Gn[n].interior.cpu().numpy() for n in pisces.TRACERS[:24]}
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), local=local.cpu().numpy(), total=total.cpu().numpy(),
             again=again.cpu().numpy(), **{"G_" + n: g for n, g in G.items()})
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_slab_stage_and_inventory_allreduce_over_nccl(tmp_path, oracle, cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    import oceanbiome_b200 as ob
    from oceanbiome_b200 import pisces, synthetic
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    # the oracle's serial sum over the GLOBAL grid (host-regenerated fields)
    full = ob.RectilinearGrid(size=SIZE, extent=EXTENT, device="cpu")
    og = oracle.Grid.like(full)
    bgc = ob.PISCES(full, scale_negatives=True)
    groups = [(m.tracers, m.scalefactors) for m in bgc.modifiers]
    names = []
    for tn, _ in groups:
        names += [t for t in tn if t not in names]
    fields = [synthetic.fill_numpy(np.zeros(og.parent_shape), og, n, *pisces.synthetic_range(n)) for n in names]
    V = float(full.dx * full.dy * full.dz[0])
    want = oracle.inventory(og, fields, oracle.make_groups(names, groups), uniform_volume=V)
    mag = oracle.inventory(og, [np.abs(f) for f in fields],
                           oracle.make_groups(names, [(t, tuple(abs(x) for x in sf)) for t, sf in groups]), uniform_volume=V)
    for r in range(world):
        assert np.all(np.abs(res[r]["total"] - want) <= 1e-12 * mag), (res[r]["total"], want)
        assert np.array_equal(res[r]["total"], res[r]["again"])        # run-to-run identical
    assert np.array_equal(res[0]["total"], res[1]["total"])            # both ranks hold the same global sums
    assert np.all(np.abs(res[0]["local"] + res[1]["local"] - want) <= 1e-12 * mag) and not np.allclose(res[0]["local"], res[1]["local"])
    # the undivided grid on one GPU: its rows are what the slabs computed, bit for bit
    grid = ob.RectilinearGrid(size=SIZE, extent=EXTENT, device=cuda)
    bgc, model = _build(ob, grid, rows=None)
    model.update_state()
    model.compute_tendencies()
    torch.cuda.synchronize()
    ny = SIZE[1] // world
    for n in pisces.TRACERS[:24]:
        whole = model.Gn[n].interior.cpu().numpy()
        for r in range(world):
            assert np.array_equal(whole[:, r * ny:(r + 1) * ny, :], res[r]["G_" + n]), (n, r)
