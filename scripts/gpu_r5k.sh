#!/bin/bash
# r5 visit k: prologue at 9 / 10 resident blocks
set -u
for rep in 1 2; do
for so in default build/variants/libobm_sn_b9.so build/variants/libobm_sn_b10.so; do
  if [ $so = default ]; then unset OBM_B200_LIB; else export OBM_B200_LIB=$PWD/$so; fi
  python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', round(d['scale_negative_calcite_fused_ms'],4))"
done
done
