#!/bin/bash
# r5 visit h: whole suite and the default bench line with the faster inventory kernel
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4.json 2> gpurun_out/bench_pisces_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_pisces_c4.json')); r=d['roofline']; i=d['inventory']
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], r['stage']['frac'], [(k['kernel'][:12], round(k['ms'],3)) for k in r['kernels']])
print('inventory', i['ms_per_stage_inventory_and_allreduce'], i['GBs_per_gpu'], i['value_with_inventory'], i['check']['ok'], i['run_to_run_identical'])"
