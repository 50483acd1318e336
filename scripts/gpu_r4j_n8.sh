#!/bin/bash
# r4 visit j (8 GPUs): the default bench line as the driver launches it at N = 8 with the r04 build — strong scaling, inventory all-reduce,
# e2e with the rows re-shared by each rank's measured host-link rate, the copies-only ceiling, weak sub-record
set -u
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4_n8_r4j.json 2> gpurun_out/bench_pisces_c4_n8_r4j.err; echo "bench n8 rc=$?"; tail -n 5 gpurun_out/bench_pisces_c4_n8_r4j.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_pisces_c4_n8_r4j.json"))
e = d["e2e"]
print(d["value"], d["ms_per_step"], d["roofline"]["stage"]["frac"], d["inventory"]["ms_per_stage_inventory_and_allreduce"], (d.get("weak") or {}).get("value"))
print(e["value"], e.get("slabs"), e["ceiling"], e["limiter"])
PY
