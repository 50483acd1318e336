#!/bin/bash
# r4 visit x: last sanity pass of the final tree — suite, smoke, default bench line, reference arm
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -n 6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 3 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_default.json 2> gpurun_out/bench_ref_default.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_default.json
