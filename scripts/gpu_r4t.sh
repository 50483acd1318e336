#!/bin/bash
# r4 visit t: the examples on the whole-run launch
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_examples.py tests/test_gpu_box_model.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 6
( time timeout 600 python examples/box.py --years 1 --out gpurun_out/box.npz ) 2>&1 | tail -n 5
( time timeout 900 python examples/data_assimilation.py ) 2>&1 | tail -n 12
