#!/bin/bash
# r5 visit d: Gⁿ of a BiogeochemicalModel in one slab (one memset per stage) — whole suite, the default bench line, the column-ensemble run timing
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -n 5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r5d.json 2> gpurun_out/bench_r5d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r5d.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], r['stage']['frac'], [(k['kernel'][:12], round(k['ms'],3)) for k in r['kernels']])"
timeout 600 python scripts/time_column_ensemble.py 2>&1 | tail -n 1 | tee gpurun_out/time_column_ensemble.json
