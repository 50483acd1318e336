#!/bin/bash
# ncu evidence for one workload: launch list of OUR kernels (shares) + one --set full capture per named kernel.
# usage (under gpurun): bash scripts/gpu_profile.sh <workload> "<kernel1> <kernel2> ..." [extra bench args]
set -u
W=${1:-lobster_c3}; KS=${2:-tendency}; shift 2 || true
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pisces_|calcite_|euphotic_|mixed_layer_|par_|scale_negative|zero_negative|npd_|carbon_sweep|inventory_|sediment_|gas_" -c 400 --csv --log-file gpurun_out/launches_$W.csv \
    python bench.py --workload $W --steps 3 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_bench_$W.log 2>&1
for K in $KS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_${W}_$K -f \
      python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_full_${W}_$K.log 2>&1
  tail -1 gpurun_out/ncu_full_${W}_$K.log
done
ls gpurun_out/
