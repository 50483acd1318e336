#!/bin/bash
# r5 visit i: resident blocks of the NPD tendency kernel (LOBSTER C3: 74 registers, 3 blocks of 256 by default) — 3 / 4 / 5 blocks
set -u
mkdir -p gpurun_out
OBM_B200_LIB=$PWD/build/variants/libobm_npd_b4.so timeout 600 python -m pytest tests/test_gpu_npd.py tests/test_gpu_box_model.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 2
for rep in 1 2 3; do
for so in default build/variants/libobm_npd_b3.so build/variants/libobm_npd_b4.so build/variants/libobm_npd_b5.so; do
  if [ $so = default ]; then unset OBM_B200_LIB; else export OBM_B200_LIB=$PWD/$so; fi
  python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', round(d['tendencies_ms'],4), round(d['tendencies_overwrite_ms'],4))"
done
done
