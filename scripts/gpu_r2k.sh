#!/bin/bash
# guard on every input being finite + all loads in one batch: PISCES parity incl. NaN / Inf inputs, one A/B timing
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pisces.py -q -m gpu 2>&1 | grep -v "^E  \|^$" | tail -14 | cut -c1-300 | tee gpurun_out/pytest_r2k.log
for lib in build/variants/libobm_before_guard.so oceanbiome.jl_b200/lib/libobm_b200.so; do
  OBM_B200_LIB=$lib python scripts/time_kernels.py pisces_c4 0.125 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['lib'][-24:], d['tendencies_ms'], d['tendencies_overwrite_ms'])" | tee -a gpurun_out/time_r2k.txt
done
