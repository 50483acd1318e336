#!/bin/bash
# r3 visit b: the GPU suite after the fixes, the TMA-staged tendency launch (A/B against the direct launch through
# OBM_PISCES_TMA, memcheck, full capture), the prologue variants with the branch-free batch of exps, the bench line.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl gpurun_out/variants.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -45 gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pisces.py -q -k "tma or fused_tendencies" -p no:cacheprovider > gpurun_out/memcheck_tma.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/memcheck_tma.log
echo "== tendency kernel: TMA-staged vs direct (16.8 M cells)"
for m in 1 0 1 0; do OBM_PISCES_TMA=$m python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('TMA=$m', *[(k, round(d[k],4)) for k in ('tendencies_ms','tendencies_overwrite_ms','scale_negative_calcite_fused_ms','light_with_column_state_ms')])" | tee -a gpurun_out/variants.txt; done
echo "== prologue variants (16.8 M cells)"
for so in build/variants/libobm_l*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in ('light_ms','light_with_column_state_ms')])" | tee -a gpurun_out/variants.txt
done
for so in build/variants/libobm_p*.so; do
  OBM_B200_LIB=$PWD/$so timeout 300 python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', *[(k, round(d[k],4)) for k in ('scale_negative_calcite_fused_ms','underlying_state_ms')])" | tee -a gpurun_out/variants.txt
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4.json 2> gpurun_out/bench_pisces_c4.err; echo "bench rc=$?"; cat gpurun_out/bench_pisces_c4.json; tail -5 gpurun_out/bench_pisces_c4.err
OBM_PISCES_TMA=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-inventory > gpurun_out/bench_pisces_c4_direct.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/bench_pisces_c4_direct.json')); print('direct launch: ms_per_step', d['ms_per_step'], [ (k['hook'], round(k['ms'],3)) for k in d['roofline']['kernels']])"
echo "== ncu"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pisces_|calcite_|par_|scale_negative|inventory_" -c 200 --csv --log-file gpurun_out/launches_pisces_c4.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for K in pisces_tendency_tma scale_negative_calcite; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/r3b_$K -f \
      python bench.py --scale 0.25 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-inventory > gpurun_out/ncu_full_$K.log 2>&1
  tail -1 gpurun_out/ncu_full_$K.log
done
ls gpurun_out/ | head -40
