"""N > 1 host logic on CPU: world-size-2 `gloo` run of the multi-GPU plumbing (oceanbiome_b200.distributed).
Each rank owns one x–y slab (`RectilinearGrid.slab`), computes its local tracer inventory (here with the oracle —
no GPU in this test), and the ONE collective of the path — the all-reduce of the per-slab inventories — must
reproduce the serial global inventory.  Slab-regenerated synthetic fields must equal the global field's rows."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import oceanbiome_b200 as ob
    import pyoracle
    from oceanbiome_b200 import distributed, synthetic
    r, w, dev = distributed.init_distributed(backend="gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    full = ob.RectilinearGrid(size=(12, 8, 5), x=(0, 12), y=(0, 16), z=(-50, 0), device="cpu")
    slab = full.slab(rank, world)
    j0, j1 = distributed.slab_ranges(full.Ny, world)[rank]
    assert (slab.Ny, slab.y) == (j1 - j0, (2.0 * j0, 2.0 * j1))
    names = ("P", "Z", "NO₃", "NH₄")
    groups = [(names, (1, 1, 1, 1)), (("P", "Z"), (6.56, 6.56))]
    og_full, og = pyoracle.Grid.like(full), pyoracle.Grid.like(slab)
    fields = []
    for n in names:  # every rank regenerates ITS rows of the global synthetic field
        g = synthetic.fill_numpy(np.zeros(og_full.parent_shape), og_full, n, *synthetic.lobster_range(n))
        loc = np.zeros(og.parent_shape)
        og.interior(loc)[...] = og_full.interior(g)[:, j0:j1, :]
        fields.append(loc)
    vol = float(slab.dx * slab.dy * slab.dz[0])
    local = pyoracle.inventory(og, fields, pyoracle.make_groups(names, groups), uniform_volume=vol)
    total = distributed.allreduce_sum(torch.from_numpy(local.copy()))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack([local, total.numpy()]))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_slab_inventory_allreduce_world_size_2(tmp_path, oracle):
    import oceanbiome_b200 as ob
    from oceanbiome_b200 import synthetic
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    # serial reference over the whole grid
    full = ob.RectilinearGrid(size=(12, 8, 5), x=(0, 12), y=(0, 16), z=(-50, 0), device="cpu")
    og = oracle.Grid.like(full)
    names = ("P", "Z", "NO₃", "NH₄")
    groups = [(names, (1, 1, 1, 1)), (("P", "Z"), (6.56, 6.56))]
    fields = [synthetic.fill_numpy(np.zeros(og.parent_shape), og, n, *synthetic.lobster_range(n)) for n in names]
    want = oracle.inventory(og, fields, oracle.make_groups(names, groups), uniform_volume=float(full.dx * full.dy * full.dz[0]))
    for r in range(world):
        np.testing.assert_allclose(res[r][1], want, rtol=1e-13)        # both ranks hold the global sum
    np.testing.assert_allclose(res[0][0] + res[1][0], want, rtol=1e-13)  # and it is the sum of the slabs
    assert not np.allclose(res[0][0], res[1][0])


def test_slab_ranges_and_errors():
    from oceanbiome_b200 import distributed
    assert distributed.slab_ranges(1024, 8) == [(128 * r, 128 * (r + 1)) for r in range(8)]
    with pytest.raises(ValueError, match="not divisible"):
        distributed.slab_ranges(10, 4)


def test_single_process_is_a_no_op(monkeypatch):
    from oceanbiome_b200 import distributed
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    rank, world, dev = distributed.init_distributed()
    assert (rank, world) == (0, 1)
    t = torch.tensor([1.0, 2.0], dtype=torch.float64)
    assert torch.equal(distributed.allreduce_sum(t.clone()), t)
