#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2b.log
tail -5 gpurun_out/pytest_gpu_r2b.log
bash scripts/sweep_variants.sh pisces_c4 0.125 light_ms tendencies_ms 2>&1 | tee gpurun_out/sweep_light_r2b.txt
bash scripts/sweep_variants.sh lobster_c3 1.0 light_ms tendencies_ms 2>&1 | tee -a gpurun_out/sweep_light_r2b.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pisces_c4_r2b.json 2> gpurun_out/bench_pisces_c4_r2b.err; cat gpurun_out/bench_pisces_c4_r2b.json; tail -3 gpurun_out/bench_pisces_c4_r2b.err
