#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2c.log
tail -5 gpurun_out/pytest_gpu_r2c.log
for E in sm dma; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --copy-engine $E > gpurun_out/bench_pisces_c4_r2c_$E.json 2> gpurun_out/bench_pisces_c4_r2c_$E.err; python -c "
import json; d=json.load(open('gpurun_out/bench_pisces_c4_r2c_$E.json')); print('$E', d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"; tail -2 gpurun_out/bench_pisces_c4_r2c_$E.err
done
