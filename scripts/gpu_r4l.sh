#!/bin/bash
# r4 visit l: the exps of a lane's four levels unclamped with one combined range test (OBM_LIGHT_EXP_BATCH) — light tests, A/B timing
set -u
mkdir -p gpurun_out
rm -f gpurun_out/variants_r4l.txt
timeout 1200 python -m pytest tests/test_gpu_light.py tests/test_gpu_pisces.py tests/test_gpu_npd.py tests/test_gpu_full_size.py tests/test_gpu_box_model.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r4l.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_r4l.log
for rep in 1 2 3; do
for so in default build/variants/libobm_eb0.so build/variants/libobm_eb1_tb3.so build/variants/libobm_eb1_mb4.so; do
  if [ $so = default ]; then unset OBM_B200_LIB; else export OBM_B200_LIB=$PWD/$so; fi
  a=$(python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['light_with_column_state_ms'],4))")
  b=$(python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['light_ms'],4))")
  echo "$so 3-band $a two-band $b" | tee -a gpurun_out/variants_r4l.txt
done
done
