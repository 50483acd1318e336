"""CPU: the parts of bench.py's contract that do not need a GPU — the reference arm (the oracle timed on the host
cores: the one place outside tests/ where the oracle may run), rank handling under torchrun, the loud failure of the
product arm without a CUDA device, and the roofline helpers."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          cwd=ROOT, timeout=600)


@pytest.mark.parametrize("workload", ["npzd_c1", "pisces_c4"])
def test_reference_arm_prints_one_json_line(workload):
    r = run_bench("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--workload", workload)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BGC tendency Gcell-updates/s" and d["unit"] == "Gcell-updates/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"] == workload and d["dtype"] == "f64" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sub-volume" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "npzd_c1",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run_bench("--steps", "1", "--workload", "npzd_c1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""  # no JSON line that could be mistaken for a measurement


def test_roofline_helpers():
    sys.path.insert(0, ROOT)
    import bench
    import json
    rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["pisces_c4"]
    traffic, note = bench.ncu_traffic("pisces_c4", 134217728 // 2)
    assert abs(traffic - rec["bytes_per_launch"] / 2) < 1 and rec["source"] in note and os.path.exists(os.path.join(ROOT, rec["source"]))
    assert 0.95 * 640 < rec["bytes_per_launch"] / rec["cells"] < 1.15 * 640  # ncu DRAM traffic within 15 % of the algorithmic bytes
    assert bench.ncu_traffic("npzd_c1", 10) == (None, None)
    f = bench.fp64_roofline("pisces_c4", 134217728, 19.03)
    assert f["unit"] == "T FP64 instr/s" and 0.4 < f["frac"] < 0.6 and 1100 < f["instr_per_cell"] < 1250
    assert bench.fp64_roofline("lobster_c3", 1, 1.0) is None
    names = set(bench.workload_table())
    assert names == {"pisces_c4", "lobster_c3", "lobster_c2", "npzd_c1", "carbon_c5"} and bench.default_workload() == "pisces_c4"
