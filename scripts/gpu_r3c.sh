#!/bin/bash
# r3 visit c: GPU suite after (1) the TMA-staged launch was dropped, (2) the prologue's cp.async staging + new solve defaults,
# (3) the rounding-faithful free iron; prologue / light variants A/B; bench line; launch list + full captures.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl gpurun_out/variants.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
echo "== variants (16.8 M cells)"
for rep in 1 2; do
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', *[(k, round(d[k],4)) for k in ('scale_negative_calcite_fused_ms','light_with_column_state_ms','tendencies_ms')])" | tee -a gpurun_out/variants.txt
for so in build/variants/libobm_*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in ('scale_negative_calcite_fused_ms','light_with_column_state_ms','tendencies_ms')])" | tee -a gpurun_out/variants.txt
done
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4.json 2> gpurun_out/bench_pisces_c4.err; echo "bench rc=$?"; cat gpurun_out/bench_pisces_c4.json; tail -5 gpurun_out/bench_pisces_c4.err
echo "== ncu"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pisces_|calcite_|par_|scale_negative|inventory_" -c 200 --csv --log-file gpurun_out/launches_pisces_c4.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for K in scale_negative_calcite; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/r3c_$K -f \
      python bench.py --scale 0.25 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-inventory > gpurun_out/ncu_full_$K.log 2>&1
  tail -1 gpurun_out/ncu_full_$K.log
done
