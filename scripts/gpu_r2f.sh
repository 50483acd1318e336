#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2f.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2f.log
tail -5 gpurun_out/pytest_gpu_r2f.log
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so python scripts/time_kernels.py pisces_c4 0.125 carbon 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', {k: round(v,4) for k,v in d.items() if k.endswith('ms') or 'ms_' in k})"
done 2>&1 | tee gpurun_out/sweep_cc_r2f.txt
