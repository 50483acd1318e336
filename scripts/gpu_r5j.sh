#!/bin/bash
# r5 visit j: NPD tendency kernels at four resident blocks — whole suite, LOBSTER / NPZD bench lines with their CPU legs, per-hook timings, box benchmark
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log
for W in lobster_c3 lobster_c2 npzd_c1; do python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; python -c "import json; d=json.load(open('gpurun_out/bench_$W.json')); r=d['roofline']; print('$W', d['value'], d['ms_per_step'], r['frac'], (r.get('stage') or {}).get('frac'), [(k['kernel'][:14], round(k['ms'],4)) for k in r.get('kernels',[])])"; done
python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | tee gpurun_out/time_kernels_lobster_c3.json
timeout 600 python scripts/time_box_model.py > gpurun_out/time_box_model.json 2> gpurun_out/time_box_model.err; python -c "
import json
for r in json.load(open('gpurun_out/time_box_model.json'))['rows']:
    if r['mode'].startswith('one'): print(r['boxes'], r['run_wall_s'], r['device_s'])"
timeout 300 python scripts/time_column_ensemble.py 2>&1 | tail -n 1 | tee gpurun_out/time_column_ensemble.json | cut -c1-400
