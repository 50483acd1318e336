#!/bin/bash
# Full evidence pass for profiles/: tests, smoke, peaks, launch lists + full captures, every bench line.
set -u
mkdir -p gpurun_out
bash scripts/gpu_check.sh pisces_c4
python scripts/stream_pattern.py > gpurun_out/stream_pattern.json 2>&1; cat gpurun_out/stream_pattern.json
python scripts/pcie_bw.py 2>&1 | tail -1 | tee gpurun_out/pcie_bw.json
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | tee gpurun_out/time_kernels_pisces_c4.json
python scripts/time_kernels.py lobster_c3 1.0 2>&1 | tail -1 | tee gpurun_out/time_kernels_lobster_c3.json
bash scripts/gpu_profile.sh pisces_c4 "pisces_tendency scale_negative_calcite par_multiband"
bash scripts/gpu_profile.sh lobster_c3 "npd_tendency par_twoband scale_negative"
for W in lobster_c3 lobster_c2 npzd_c1 carbon_c5; do
  python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; cat gpurun_out/bench_$W.json
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_pisces_c4.json 2> gpurun_out/bench_ref_pisces_c4.err; cat gpurun_out/bench_ref_pisces_c4.json
