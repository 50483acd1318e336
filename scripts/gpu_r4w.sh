#!/bin/bash
# r4 visit w: block shape of the PISCES tendency kernel once more (128x3 default, 64x6, 32x12), four alternating passes at 1/8 size and one at full size
set -u
mkdir -p gpurun_out
rm -f gpurun_out/variants_r4w.txt
OBM_B200_LIB=$PWD/build/variants/libobm_pb64.so timeout 600 python -m pytest tests/test_gpu_pisces.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 2
for rep in 1 2 3 4; do
for so in default build/variants/libobm_pb64.so build/variants/libobm_pb32.so; do
  if [ $so = default ]; then unset OBM_B200_LIB; else export OBM_B200_LIB=$PWD/$so; fi
  python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', round(d['tendencies_ms'],4), round(d['tendencies_overwrite_ms'],4))" | tee -a gpurun_out/variants_r4w.txt
done
done
for so in default build/variants/libobm_pb64.so default build/variants/libobm_pb64.so; do
  if [ $so = default ]; then unset OBM_B200_LIB; else export OBM_B200_LIB=$PWD/$so; fi
  python scripts/time_kernels.py pisces_c4 1.0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('full size $so', round(d['tendencies_ms'],4), round(d['tendencies_overwrite_ms'],4))" | tee -a gpurun_out/variants_r4w.txt
done
