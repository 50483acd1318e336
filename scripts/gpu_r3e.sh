#!/bin/bash
# r3 visit e: GPU suite with the Ω solve's error extrapolation, the clamped branch-free exp in the PAR scans and the
# level-table-free prologue; exit-threshold variants; smoke; bench lines (C4 default, C3); launch list; full captures.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl gpurun_out/variants_r3e.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
for rep in 1 2; do
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', *[(k, round(d[k],4)) for k in ('scale_negative_calcite_fused_ms','light_with_column_state_ms','tendencies_ms')])" | tee -a gpurun_out/variants_r3e.txt
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in ('scale_negative_calcite_fused_ms','light_with_column_state_ms','tendencies_ms')])" | tee -a gpurun_out/variants_r3e.txt
done
done
for so in build/variants/libobm_q_t5e5.so; do
  OBM_B200_LIB=$PWD/$so timeout 900 python -m pytest tests/test_gpu_carbon.py tests/test_gpu_pisces.py tests/test_gpu_full_size.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_t5e5.log 2>&1; echo "pytest($so) rc=$?"; tail -5 gpurun_out/pytest_t5e5.log
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4.json 2> gpurun_out/bench_pisces_c4.err; echo "bench rc=$?"; cat gpurun_out/bench_pisces_c4.json; tail -5 gpurun_out/bench_pisces_c4.err
timeout 600 python bench.py --workload lobster_c3 --steps 10 --warmup 3 > gpurun_out/bench_lobster_c3.json 2> gpurun_out/bench_lobster_c3.err; echo "bench c3 rc=$?"; tail -3 gpurun_out/bench_lobster_c3.err
echo "== ncu"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pisces_|calcite_|par_|scale_negative|inventory_" -c 200 --csv --log-file gpurun_out/launches_pisces_c4.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for K in scale_negative_calcite par_multiband pisces_tendency; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/r3e_$K -f \
      python bench.py --scale 0.25 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-inventory > gpurun_out/ncu_full_$K.log 2>&1
  tail -1 gpurun_out/ncu_full_$K.log
done
