#!/bin/bash
# Time the PISCES tendency kernel for every library variant under build/variants (gpurun helper).
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', round(d['tendencies_ms'],3))"
done
