#!/bin/bash
# memcheck over the parameter-ensemble / NPD / box-model tests, then the round-end sequence (gpu_final.sh)
set -u
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck_r2h.log python -m pytest tests/test_gpu_parameter_ensemble.py tests/test_gpu_npd.py tests/test_gpu_box_model.py tests/test_gpu_examples.py -x -q -m gpu > gpurun_out/sanitizer_memcheck_r2h_pytest.log 2>&1
echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck_r2h.log; tail -3 gpurun_out/sanitizer_memcheck_r2h_pytest.log; grep "ERROR SUMMARY" gpurun_out/sanitizer_memcheck_r2h.log | sort | uniq -c | head -3
bash scripts/gpu_final.sh
