#!/bin/bash
# usage: sweep_env.sh VAR v1 v2 ... — time the PISCES kernels with VAR set to each value (gpurun helper)
var=$1; shift
for v in "$@"; do
  env $var=$v python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$var=$v', round(d['tendencies_ms'],3))"
done
