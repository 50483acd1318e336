#!/bin/bash
# r4 visit k: lean batched log in the PAR scans, the solubility product's exp in the constants' batch — GPU suite with the default build, A/B timing
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl gpurun_out/variants_r4k.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
K="scale_negative_calcite_fused_ms light_with_column_state_ms tendencies_ms"
for rep in 1 2 3; do
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', *[(k, round(d[k],4)) for k in '$K'.split()])" | tee -a gpurun_out/variants_r4k.txt
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in '$K'.split()])" | tee -a gpurun_out/variants_r4k.txt
done
done
for rep in 1 2; do
python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lobster default', round(d['light_ms'],4))" | tee -a gpurun_out/variants_r4k.txt
OBM_B200_LIB=$PWD/build/variants/libobm_llog0.so python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lobster llog0', round(d['light_ms'],4))" | tee -a gpurun_out/variants_r4k.txt
done
