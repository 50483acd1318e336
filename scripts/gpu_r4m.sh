#!/bin/bash
# r4 visit m: full captures of the prologue and the 3-band scan of the current build at 1/8 size (for the per-opcode reading); negs tests
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_negs.py tests/test_gpu_pisces.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r4m.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_r4m.log
for K in scale_negative_calcite par_multiband; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/r4m_$K -f \
      python bench.py --scale 0.125 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-inventory > gpurun_out/ncu_full_r4m_$K.log 2>&1
  tail -n 1 gpurun_out/ncu_full_r4m_$K.log
done
