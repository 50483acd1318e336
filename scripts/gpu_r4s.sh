#!/bin/bash
# r4 visit s: NPD family without multiply-add contraction (same bits from every launch form) — the whole GPU suite, LOBSTER kernel timings, the box benchmark
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 3 gpurun_out/smoke.log
for rep in 1 2; do python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lobster_c3', d['tendencies_ms'], d['tendencies_overwrite_ms'])"; done
for W in lobster_c3 lobster_c2 npzd_c1; do python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; python -c "import json; d=json.load(open('gpurun_out/bench_$W.json')); print('$W', d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
timeout 900 python scripts/time_box_model.py > gpurun_out/time_box_model.json 2> gpurun_out/time_box_model.err; echo "box rc=$?"; tail -n 3 gpurun_out/time_box_model.err
python - <<'PY'
import json
for r in json.load(open("gpurun_out/time_box_model.json"))["rows"]:
    print(r["boxes"], r["fused_tendency_and_substep"], r["mode"][:10], r["run_wall_s"], r["device_s"], r["us_per_rk3_stage"], r["P_end_member0"])
PY
