#!/bin/bash
# r4 visit p (4 GPUs): the default bench line as the driver launches it at N = 4 with the r04 build
set -u
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4_n4_r4p.json 2> gpurun_out/bench_pisces_c4_n4_r4p.err; echo "bench n4 rc=$?"; tail -n 3 gpurun_out/bench_pisces_c4_n4_r4p.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_pisces_c4_n4_r4p.json"))
e = d["e2e"]
print(d["value"], d["ms_per_step"], d["roofline"]["stage"]["frac"], d["inventory"]["ms_per_stage_inventory_and_allreduce"], (d.get("weak") or {}).get("value"), d["clocks"])
print(e["value"], e.get("slabs"), e["ceiling"])
PY
