#!/bin/bash
# r4 visit q: the whole run of a box-model ensemble in one launch (obm_npd_box_run) — box-model tests, the reference's box benchmark
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_box_model.py tests/test_gpu_npd.py tests/test_gpu_parameter_ensemble.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r4q.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_r4q.log
timeout 900 python scripts/time_box_model.py > gpurun_out/time_box_model.json 2> gpurun_out/time_box_model.err; echo "box rc=$?"; tail -n 3 gpurun_out/time_box_model.err
python - <<'PY'
import json
for r in json.load(open("gpurun_out/time_box_model.json"))["rows"]:
    print(r["boxes"], r["fused_tendency_and_substep"], r["mode"][:10], r["run_wall_s"], r["device_s"], r["us_per_rk3_stage"], r["P_end_member0"])
PY
