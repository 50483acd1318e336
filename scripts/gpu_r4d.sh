#!/bin/bash
# r4 visit d: two-band PAR scan with four levels per lane (default now) — light / NPD / box-model tests, timing against the
# one-level-per-lane form and two launch-bound variants on lobster_c3; the e2e probe on one GPU (copy pattern at full size)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/variants_r4d.txt
timeout 1200 python -m pytest tests/test_gpu_light.py tests/test_gpu_npd.py tests/test_gpu_full_size.py tests/test_gpu_box_model.py tests/test_gpu_host_stage.py tests/test_gpu_examples.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_r4d.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_r4d.log
for V in tb_b4 tb_b5; do
OBM_B200_LIB=$PWD/build/variants/libobm_$V.so timeout 600 python -m pytest tests/test_gpu_light.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_$V.log 2>&1; echo "pytest($V) rc=$?"; tail -n 2 gpurun_out/pytest_$V.log
done
K="scale_negative_ms light_ms tendencies_ms"
for rep in 1 2; do
python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', *[(k, round(d[k],4)) for k in '$K'.split()])" | tee -a gpurun_out/variants_r4d.txt
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in '$K'.split()])" | tee -a gpurun_out/variants_r4d.txt
done
done
timeout 600 python scripts/e2e_probe.py dma:16 copies:16 flat:16 flat:1 h2d:16 d2h:16 copies:4 copies:64 dma:16 > gpurun_out/e2e_probe_n1.jsonl 2> gpurun_out/e2e_probe_n1.err; echo "probe rc=$?"; cat gpurun_out/e2e_probe_n1.jsonl; tail -n 3 gpurun_out/e2e_probe_n1.err
