#!/bin/bash
# r4 visit n: the round's final evidence pass on one GPU with the r04 build — suite, smoke, peaks, launch lists + full captures, every bench line,
# reference arm, sanitizers, box-model ensemble timing
set -u
bash scripts/gpu_round.sh
bash scripts/gpu_sanitize.sh
timeout 900 python scripts/time_box_model.py > gpurun_out/time_box_model.json 2> gpurun_out/time_box_model.err; echo "box rc=$?"
