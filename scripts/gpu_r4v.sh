#!/bin/bash
# r4 visit v: LOBSTER C3 evidence with the final build (NPD family without multiply-add contraction): bench line with its CPU legs, launch list, full captures
set -u
mkdir -p gpurun_out
python bench.py --workload lobster_c3 --steps 10 --warmup 3 > gpurun_out/bench_lobster_c3.json 2> gpurun_out/bench_lobster_c3.err; cut -c1-300 gpurun_out/bench_lobster_c3.json
python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | tee gpurun_out/time_kernels_lobster_c3.json
bash scripts/gpu_profile.sh lobster_c3 "npd_tendency par_twoband scale_negative"
for W in lobster_c2 npzd_c1; do python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; cut -c1-200 gpurun_out/bench_$W.json; done
