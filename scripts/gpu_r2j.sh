#!/bin/bash
# NaN-input guard as a data dependency (all input loads in one batch): parity, NaN patterns, A/B timing on one box
set -u
mkdir -p gpurun_out
OLD=build/variants/libobm_before_guard.so
python -m pytest tests/test_gpu_pisces.py -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_r2j_new.log
OBM_B200_LIB=$OLD python -m pytest tests/test_gpu_pisces.py -q -m gpu -k nan_inputs 2>&1 | tail -6 | tee gpurun_out/pytest_r2j_old.log
for lib in $OLD oceanbiome.jl_b200/lib/libobm_b200.so $OLD oceanbiome.jl_b200/lib/libobm_b200.so; do
  OBM_B200_LIB=$lib python scripts/time_kernels.py pisces_c4 0.125 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['lib'][-24:], d['tendencies_ms'], d['tendencies_overwrite_ms'])" | tee -a gpurun_out/time_r2j.txt
done
