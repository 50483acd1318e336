#!/bin/bash
# r3 visit j: the round's evidence pass on one GPU — suite, smoke, sanitizers, peaks, launch lists + full captures, every bench line, reference arm
set -u
bash scripts/gpu_round.sh
bash scripts/gpu_sanitize.sh
timeout 900 python scripts/time_box_model.py > gpurun_out/time_box_model.json 2> gpurun_out/time_box_model.err; echo "box rc=$?"
