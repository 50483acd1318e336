#!/bin/bash
# r4 visit y: FP32 pre-solve as fused multiply-add chains, no extrapolation block after a single FP64 step — parity, A/B against the previous header
set -u
mkdir -p gpurun_out
rm -f gpurun_out/variants_r4y.txt
timeout 1200 python -m pytest tests/test_gpu_carbon.py tests/test_gpu_pisces.py tests/test_gpu_full_size.py tests/test_gpu_gas_exchange.py tests/test_gpu_negs.py tests/test_gpu_host_stage.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 3
for rep in 1 2 3; do
for so in default build/variants/libobm_pre_old.so; do
  if [ $so = default ]; then unset OBM_B200_LIB; else export OBM_B200_LIB=$PWD/$so; fi
  python scripts/time_kernels.py pisces_c4 0.125 carbon 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', round(d['scale_negative_calcite_fused_ms'],4), 'carbon sweep 25M', round(d['carbon_sweep_ms_25M'],4))" | tee -a gpurun_out/variants_r4y.txt
done
done
