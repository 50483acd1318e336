#!/bin/bash
# One GPU visit: parity tests, smoke, DFMA peak, a bench line, and the ncu launch list.
# usage (under gpurun): bash scripts/gpu_check.sh [workload]
set -u
W=${1:-lobster_c3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python - > gpurun_out/dfma.log 2>&1 <<'PY'
import torch, oceanbiome_b200 as ob
lib = ob.load_library()
s = torch.empty(148*8*256, dtype=torch.float64, device="cuda")
for it in (4096, 16384, 16384):
    print("DFMA/s", lib.obm_fp64_peak_dfma_per_s(s.data_ptr(), it, None))
PY
cat gpurun_out/dfma.log
python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; cat gpurun_out/bench_$W.json; tail -3 gpurun_out/bench_$W.err
