#!/bin/bash
# r4 visit o: the e2e stage with its H2D sources in write-combined pinned memory (one GPU, full size)
set -u
mkdir -p gpurun_out
timeout 600 python scripts/e2e_probe.py dma:16 wc:16 dma:16 wc:16 > gpurun_out/e2e_probe_wc.jsonl 2> gpurun_out/e2e_probe_wc.err; echo "probe rc=$?"; cat gpurun_out/e2e_probe_wc.jsonl; tail -n 3 gpurun_out/e2e_probe_wc.err
