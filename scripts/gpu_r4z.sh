#!/bin/bash
# r4 visit z: PISCES C4 and carbonate-sweep evidence with the final build — whole suite, smoke, bench lines with their CPU legs, per-hook timings,
# launch list and full captures of the three PISCES kernels
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -n 5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 3 gpurun_out/smoke.log
python bench.py --workload pisces_c4 --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4.json 2> gpurun_out/bench_pisces_c4.err; cut -c1-200 gpurun_out/bench_pisces_c4.json
python bench.py --workload carbon_c5 --steps 10 --warmup 3 > gpurun_out/bench_carbon_c5.json 2> gpurun_out/bench_carbon_c5.err; cut -c1-200 gpurun_out/bench_carbon_c5.json
python scripts/time_kernels.py pisces_c4 0.125 2>&1 | tail -1 | tee gpurun_out/time_kernels_pisces_c4.json
bash scripts/gpu_profile.sh pisces_c4 "pisces_tendency scale_negative_calcite par_multiband"
