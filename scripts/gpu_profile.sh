#!/bin/bash
# ncu evidence for one workload: launch list (shares) + one --set full capture of the top kernels.
# usage (under gpurun): bash scripts/gpu_profile.sh <workload> <kernel-regex> [extra bench args]
set -u
W=${1:-lobster_c3}; K=${2:-tendency}; shift 2 || true
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$W.csv \
    python bench.py --workload $W --steps 3 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_bench_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 2 -o gpurun_out/prof_$W -f \
    python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_full_$W.log 2>&1
tail -2 gpurun_out/ncu_full_$W.log
ls -la gpurun_out/
