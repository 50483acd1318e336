#!/bin/bash
# r4 visit b (8 GPUs): the e2e probe at N = 8 — what limits the host-staged stage when eight ranks copy at once
set -u
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 scripts/e2e_probe.py dma:8 dma:1 dma:2 dma:4 dma:16 copies:8 copies:1 h2d:8 d2h:8 flat:8 flat:1 sm:8 dma:8 > gpurun_out/e2e_probe_n8.jsonl 2> gpurun_out/e2e_probe_n8.err; echo "probe rc=$?"
cat gpurun_out/e2e_probe_n8.jsonl; tail -5 gpurun_out/e2e_probe_n8.err
