#!/bin/bash
# r4 visit i (2 GPUs): the e2e leg's re-sharing branch forced (threshold 1.0) so that it is exercised at N = 2 as it will run at N = 8
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 --no-weak --no-cpu-baseline --e2e-balance-threshold 1.0 > gpurun_out/bench_pisces_c4_n2_forced.json 2> gpurun_out/bench_pisces_c4_n2_forced.err; echo "bench n2 rc=$?"; tail -n 5 gpurun_out/bench_pisces_c4_n2_forced.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_pisces_c4_n2_forced.json"))
e = d["e2e"]
print(d["value"], e["value"], e.get("slabs"), e["h2d_bytes_per_step"], e["ceiling"]["frac"])
PY
