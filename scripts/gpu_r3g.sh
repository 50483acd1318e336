#!/bin/bash
# r3 visit g: GPU suite (exp clamp fixed, fused step fixed); table-exp variants: parity with the variant libraries, timing A/B
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl gpurun_out/variants_r3g.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
OBM_B200_LIB=$PWD/build/variants/libobm_e3_table.so timeout 900 python -m pytest tests/test_gpu_pisces.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_e3.log 2>&1; echo "pytest(e3) rc=$?"; tail -4 gpurun_out/pytest_e3.log
OBM_B200_LIB=$PWD/build/variants/libobm_l4_table.so timeout 900 python -m pytest tests/test_gpu_light.py tests/test_gpu_pisces.py tests/test_gpu_full_size.py tests/test_gpu_negs.py tests/test_gpu_sinking.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_l4.log 2>&1; echo "pytest(l4) rc=$?"; tail -4 gpurun_out/pytest_l4.log
OBM_B200_LIB=$PWD/build/variants/libobm_c3_table.so timeout 900 python -m pytest tests/test_gpu_carbon.py tests/test_gpu_pisces.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_c3.log 2>&1; echo "pytest(c3) rc=$?"; tail -4 gpurun_out/pytest_c3.log
for rep in 1 2; do
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', *[(k, round(d[k],4)) for k in ('scale_negative_calcite_fused_ms','light_with_column_state_ms','tendencies_ms','tendencies_overwrite_ms')])" | tee -a gpurun_out/variants_r3g.txt
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in ('scale_negative_calcite_fused_ms','light_with_column_state_ms','tendencies_ms','tendencies_overwrite_ms')])" | tee -a gpurun_out/variants_r3g.txt
done
done
OBM_B200_LIB=$PWD/build/variants/libobm_e3_table.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:pisces_tendency -s 3 -c 1 -o gpurun_out/r3g_e3 -f \
      python bench.py --scale 0.25 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-inventory > gpurun_out/ncu_full_e3.log 2>&1
tail -1 gpurun_out/ncu_full_e3.log
