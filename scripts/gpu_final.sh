#!/bin/bash
# What the driver runs at round end, in one visit: GPU tests, smoke, the default bench line, the reference arm.
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log; tail -4 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-300 gpurun_out/bench_final_ref.json
SECONDS=0; python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-400 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err; echo "bench wall ${SECONDS} s"
