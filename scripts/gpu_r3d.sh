#!/bin/bash
# r3 visit d: software-pipelined persistent tendency launch — parity (PISCES GPU tests with the variant library), memcheck, timing A/B
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
rm -f gpurun_out/variants_r3d.txt
for so in build/variants/libobm_t_pipe3.so; do
  OBM_B200_LIB=$PWD/$so timeout 900 python -m pytest tests/test_gpu_pisces.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_pipe.log 2>&1; echo "pytest($so) rc=$?"; tail -5 gpurun_out/pytest_pipe.log
  OBM_B200_LIB=$PWD/$so timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pisces.py -q -k "fused_tendencies" -p no:cacheprovider > gpurun_out/memcheck_pipe.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_pipe.log
done
for rep in 1 2; do
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', *[(k, round(d[k],4)) for k in ('scale_negative_calcite_fused_ms','light_with_column_state_ms','tendencies_ms','tendencies_overwrite_ms')])" | tee -a gpurun_out/variants_r3d.txt
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in ('scale_negative_calcite_fused_ms','light_with_column_state_ms','tendencies_ms','tendencies_overwrite_ms')])" | tee -a gpurun_out/variants_r3d.txt
done
done
OBM_B200_LIB=$PWD/build/variants/libobm_t_pipe3.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:pisces_tendency -s 3 -c 1 -o gpurun_out/r3d_pipe -f \
      python bench.py --scale 0.25 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-inventory > gpurun_out/ncu_full_pipe.log 2>&1
tail -1 gpurun_out/ncu_full_pipe.log
