#!/bin/bash
# r4 visit a (2 GPUs): the e2e probe (what limits the host-staged stage when several ranks copy at once) — script check at N = 2
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/e2e_probe.py > gpurun_out/e2e_probe_n2.jsonl 2> gpurun_out/e2e_probe_n2.err; echo "probe rc=$?"
cat gpurun_out/e2e_probe_n2.jsonl; tail -5 gpurun_out/e2e_probe_n2.err
