#!/bin/bash
# parameter-sweep ensembles: parity tests + the reference's box-model benchmark as a device ensemble
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parameter_ensemble.py tests/test_gpu_box_model.py tests/test_gpu_npd.py tests/test_gpu_examples.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_r2g.log
timeout 300 python scripts/time_box_model.py > gpurun_out/time_box_model.json 2> gpurun_out/time_box_model.err; echo "time_box_model exit $?"; tail -5 gpurun_out/time_box_model.err
cat gpurun_out/time_box_model.json | head -80
