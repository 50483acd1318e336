#!/bin/bash
# r4 visit g (2 GPUs): the default bench line as the driver launches it at N = 2 — strong scaling, inventory all-reduce, e2e with the rows
# re-shared by each rank's measured host-link rate and the copies-only ceiling; the 2-rank NCCL test
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4_n2_r4g.json 2> gpurun_out/bench_pisces_c4_n2_r4g.err; echo "bench n2 rc=$?"; tail -n 5 gpurun_out/bench_pisces_c4_n2_r4g.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_pisces_c4_n2_r4g.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["stage"]["frac"])
print(json.dumps(d["e2e"], indent=1)[:3000])
PY
