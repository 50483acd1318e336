#!/bin/bash
# r4 visit c: GPU suite with the FP32 pre-solve in the carbonate Newton iteration; parity of the 4-levels-per-lane PAR scan variants;
# timing A/B of every variant (two passes, alternating)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl gpurun_out/variants_r4c.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for V in s4 s4_b4; do
OBM_B200_LIB=$PWD/build/variants/libobm_$V.so timeout 900 python -m pytest tests/test_gpu_light.py tests/test_gpu_pisces.py tests/test_gpu_full_size.py tests/test_gpu_host_stage.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_$V.log 2>&1; echo "pytest($V) rc=$?"; tail -4 gpurun_out/pytest_$V.log
done
K="scale_negative_calcite_fused_ms light_with_column_state_ms tendencies_ms tendencies_overwrite_ms"
for rep in 1 2; do
python scripts/time_kernels.py pisces_c4 0.125 carbon 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', *[(k, round(d[k],4)) for k in '$K carbon_sweep_ms_25M'.split()])" | tee -a gpurun_out/variants_r4c.txt
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in '$K'.split()])" | tee -a gpurun_out/variants_r4c.txt
done
done
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4_r4c.json 2> gpurun_out/bench_pisces_c4_r4c.err; cut -c1-1500 gpurun_out/bench_pisces_c4_r4c.json; tail -3 gpurun_out/bench_pisces_c4_r4c.err
