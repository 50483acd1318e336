#!/bin/bash
# r5 visit g: inventory kernel templated on the number of groups (46 registers for PISCES' five budgets) at 4 / 5 / 6 / 8 blocks per SM
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_negs.py tests/test_gpu_full_size.py tests/test_gpu_distributed.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 3
for rep in 1 2; do
for so in default build/variants/libobm_inv_sm5.so build/variants/libobm_inv_sm6.so build/variants/libobm_inv_sm8.so; do
  if [ $so = default ]; then unset OBM_B200_LIB; else export OBM_B200_LIB=$PWD/$so; fi
  python scripts/time_inventory.py 0.25 2>&1 | tail -1 | cut -c1-120
done
done
