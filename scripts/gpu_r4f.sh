#!/bin/bash
# r4 visit f: block shape of the PISCES tendency kernel (threads x resident blocks at 168 registers), series for ln(1 - 0.001005 S) in the prologue
set -u
mkdir -p gpurun_out
rm -f gpurun_out/variants_r4f.txt
OBM_B200_LIB=$PWD/build/variants/libobm_c_log1m.so timeout 900 python -m pytest tests/test_gpu_carbon.py tests/test_gpu_pisces.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_c_log1m.log 2>&1; echo "pytest(c_log1m) rc=$?"; tail -n 3 gpurun_out/pytest_c_log1m.log
OBM_B200_LIB=$PWD/build/variants/libobm_pb192.so timeout 900 python -m pytest tests/test_gpu_pisces.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_pb192.log 2>&1; echo "pytest(pb192) rc=$?"; tail -n 3 gpurun_out/pytest_pb192.log
K="scale_negative_calcite_fused_ms light_with_column_state_ms tendencies_ms tendencies_overwrite_ms"
for rep in 1 2; do
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', *[(k, round(d[k],4)) for k in '$K'.split()])" | tee -a gpurun_out/variants_r4f.txt
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in '$K'.split()])" | tee -a gpurun_out/variants_r4f.txt
done
done
