#!/bin/bash
# r5 visit f: inventory kernel with four tracers' loads in flight — distributed / negs tests, A/B timing (batch 1 / 4 / 8, 4 or 8 blocks per SM)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_negs.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 3
for rep in 1 2; do
for so in default build/variants/libobm_inv_b1.so build/variants/libobm_inv_b8.so build/variants/libobm_inv_b4_sm8.so; do
  if [ $so = default ]; then unset OBM_B200_LIB; else export OBM_B200_LIB=$PWD/$so; fi
  python scripts/time_inventory.py 0.25 2>&1 | tail -1 | cut -c1-200
done
done
