#!/bin/bash
# Time the hot kernels for every library variant under build/variants (gpurun helper).
# usage: sweep_variants.sh [workload] [scale] [json key ...]
W=${1:-pisces_c4}; S=${2:-0.125}; shift 2 || true
KEYS=${*:-tendencies_ms}
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so python scripts/time_kernels.py $W $S 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in '$KEYS'.split()])"
done
