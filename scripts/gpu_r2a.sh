#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pisces.py tests/test_gpu_negs.py tests/test_gpu_carbon.py -x -q -m gpu > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2a.log
tail -8 gpurun_out/pytest_gpu_r2a.log
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | tee gpurun_out/time_kernels_pisces_c4_r2a.json
bash scripts/sweep_variants.sh 2>&1 | tee gpurun_out/sweep_variants_r2a.txt
python scripts/pcie_bw.py 2>&1 | tail -1 | tee gpurun_out/pcie_bw.json
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_pisces_c4_r2a.json 2> gpurun_out/bench_pisces_c4_r2a.err; cat gpurun_out/bench_pisces_c4_r2a.json; tail -3 gpurun_out/bench_pisces_c4_r2a.err
