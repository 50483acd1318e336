#!/bin/bash
# r4 visit h: per-level tables of the Ω solve in global memory (OBM_SN_LEVEL=2: built by a 3 µs launch, no shared memory, no barrier)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/variants_r4h.txt
OBM_B200_LIB=$PWD/build/variants/libobm_lvl2.so timeout 900 python -m pytest tests/test_gpu_carbon.py tests/test_gpu_pisces.py tests/test_gpu_full_size.py tests/test_gpu_negs.py tests/test_gpu_host_stage.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_lvl2.log 2>&1; echo "pytest(lvl2) rc=$?"; tail -n 3 gpurun_out/pytest_lvl2.log
K="scale_negative_calcite_fused_ms light_with_column_state_ms tendencies_ms tendencies_overwrite_ms"
for rep in 1 2 3; do
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', *[(k, round(d[k],4)) for k in '$K'.split()])" | tee -a gpurun_out/variants_r4h.txt
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in '$K'.split()])" | tee -a gpurun_out/variants_r4h.txt
done
done
