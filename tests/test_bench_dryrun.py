"""CPU: bench.py's own control flow — workload construction, the marked stage, per-kernel and stage roofline assembly, the
inventory leg with its oracle check, the JSON line — exercised WITHOUT a GPU: the C library is replaced by the call
recorder of test_host_orchestration.py (no kernel runs), CUDA events by a stand-in clock.  Numbers are meaningless here;
what is checked is that every key of the contract is present, that both arms emit the SAME `config`, and that the
sharding arithmetic of a multi-rank run is right.  (The e2e leg needs real streams and is covered on the GPU box.)"""
import importlib
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from test_host_orchestration import MODULES, Recorder  # noqa: E402
from oceanbiome_b200 import _lib  # noqa: E402


class FakeEvent:
    clock = 0.0

    def __init__(self, enable_timing=True):
        self.t = None

    def record(self):
        FakeEvent.clock += 1.0
        self.t = FakeEvent.clock

    def elapsed_time(self, other):
        return other.t - self.t


@pytest.fixture
def dry(monkeypatch):
    rec = Recorder()
    monkeypatch.setattr(_lib, "load", lambda path=None: rec)
    for m in MODULES + ["distributed"]:
        mod = importlib.import_module(f"oceanbiome_b200.{m}")
        if hasattr(mod, "require_cuda"):
            monkeypatch.setattr(mod, "require_cuda", lambda *a, **k: None)
        if hasattr(mod, "current_stream_ptr"):
            monkeypatch.setattr(mod, "current_stream_ptr", lambda device: None)
    import oceanbiome_b200 as ob
    from oceanbiome_b200 import distributed
    monkeypatch.setattr(ob, "load_library", lambda *a, **k: rec)
    monkeypatch.setattr(distributed, "init_distributed", lambda backend=None: (0, 1, torch.device("cpu")))
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    return rec


def run_bench(capfd, argv):
    import bench
    old = sys.argv
    sys.argv = ["bench.py"] + argv
    try:
        bench.main()
    finally:
        sys.argv = old
    out = capfd.readouterr().out.strip().splitlines()
    return json.loads(out[-1])


@pytest.mark.parametrize("workload,scale", [("pisces_c4", 0.004), ("lobster_c3", 0.008), ("npzd_c1", 1.0)])
def test_bench_line_has_every_key_of_the_contract(dry, capfd, workload, scale):
    line = run_bench(capfd, ["--workload", workload, "--scale", str(scale), "--steps", "2", "--no-e2e", "--graph", "off", "--no-cpu-baseline"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "gpu_launches", "clocks", "roofline", "e2e", "cpu_baseline", "inventory"):
        assert key in line, key
    assert line["warmup"] == 3 and line["n_gpus"] == 1 and line["dtype"] == "f64" and line["vs_baseline"] is None
    r = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "stage", "kernels", "kernel_ms"):
        assert key in r, key
    hooks = [k["hook"] for k in r["kernels"]]
    assert hooks[:2] == ["modifiers", "light"] and "tendencies" in hooks
    if workload == "lobster_c3":
        assert "sediment" in hooks and "sediment_tendencies" in hooks  # configs[2] as named: the sediment launches are in the step
        assert "obm_sediment_update_state" in dry.calls and "obm_sediment_update_tendencies" in dry.calls
    assert abs(sum(k["algorithmic_bytes_per_cell"] for k in r["kernels"]) - r["stage"]["algorithmic_bytes_per_cell"]) < 1e-9
    if workload == "pisces_c4":
        assert [k["algorithmic_bytes_per_cell"] for k in r["kernels"]] == [192, 48, 640] and r["stage"]["algorithmic_bytes_per_cell"] == 880
        assert line["gpu_launches"] == 3 * 2  # three launches per stage
        assert dry.calls.count("obm_inventory") >= 2 and line["inventory"]["groups"] == 5 and line["inventory"]["tracers_read"] == 20
        assert line["inventory"]["check"]["cells"] > 0  # the recorder computes nothing, so only the plumbing is checked here


def test_both_arms_emit_the_same_config(dry, capfd):
    import bench
    mine = run_bench(capfd, ["--scale", "0.004", "--steps", "2", "--no-e2e", "--graph", "off", "--no-cpu-baseline", "--no-inventory"])
    assert mine["config"] == bench.config_dict("pisces_c4", 1, "strong")
    for world in (1, 2, 4, 8):
        c = bench.config_dict("pisces_c4", world, "strong")
        assert c["scaling"] == "strong" and ("slabs" in c["parallelism"]) == (world > 1)
        assert bench.shards("pisces_c4", world, "strong") == ("strong", 1024 // world if world > 1 else None)
    assert bench.shards("lobster_c2", 4, "strong")[0] == "weak" and bench.shards("carbon_c5", 2, "strong")[0] == "weak"
    assert bench.reference_budget("pisces_c4") * 64 >= 1024 * 1024 * 128  # the reference arm's sample is ≥ 1/64 of the grid


def test_slab_rows_are_the_global_rows(dry):
    import bench
    from oceanbiome_b200 import synthetic
    whole = bench.Workload("lobster_c3", torch.device("cpu"), 16 / 512)
    assert whole.grid.Ny == 16 and whole.grid.dz[0] > 1.9 * whole.grid.dz[-1]  # Eady stretching
    half = bench.Workload("lobster_c3", torch.device("cpu"), 1.0, rows=(8, 8, 16))
    assert half.grid.Ny == 8
    for n in ("NO₃", "P", "DIC"):
        assert torch.equal(half.model.tracers[n].interior, whole.model.tracers[n].interior[:, 8:16, :]), n
