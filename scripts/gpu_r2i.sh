#!/bin/bash
# last visit of the round: the whole GPU suite, smoke, and the default bench line without the e2e leg (unchanged since gpu_r2h)
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_last.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_last.log; tail -3 gpurun_out/pytest_gpu_last.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SECONDS=0; timeout 100 python bench.py --no-e2e > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench exit $? wall ${SECONDS} s"; python -c "
import json; d=json.load(open('gpurun_out/bench_last.json')); print(d['value'], d['clocks']['sm_mhz'], d['cpu_baseline'])"
