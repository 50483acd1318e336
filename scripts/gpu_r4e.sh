#!/bin/bash
# r4 visit e: GPU suite with obm_pisces_tendencies_rows (ModelLatitude) and the two-band scan at 4 blocks; smoke; the default bench line
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 3 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4_r4e.json 2> gpurun_out/bench_pisces_c4_r4e.err; cut -c1-600 gpurun_out/bench_pisces_c4_r4e.json; tail -n 3 gpurun_out/bench_pisces_c4_r4e.err
