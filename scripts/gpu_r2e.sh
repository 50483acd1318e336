#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2e.log
tail -5 gpurun_out/pytest_gpu_r2e.log
python scripts/time_e2e.py pisces_c4 0.5 dma:16 numa dma:16 dma:32 2>&1 | grep -v Warn | tee gpurun_out/time_e2e_r2e.txt
