#!/bin/bash
# r4 visit u: compute-sanitizer over the kernels added after the evidence pass — the whole-run box launch (shared-memory staging, stage coefficients
# rewritten between barriers) and the per-row PISCES launch (shared-memory copy of the kernel arguments): racecheck + memcheck
set -u
mkdir -p gpurun_out
T="tests/test_gpu_box_model.py::test_whole_run_launch_is_bit_identical_to_the_replayed_graph tests/test_gpu_pisces.py::test_model_latitude_rows_match_oracle_row_by_row"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitizer_racecheck_new.log python -m pytest $T -x -q -m gpu > gpurun_out/sanitizer_racecheck_new_pytest.log 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/sanitizer_racecheck_new.log; tail -n 3 gpurun_out/sanitizer_racecheck_new_pytest.log; grep "RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck_new.log | sort | uniq -c | head
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck_new.log python -m pytest $T tests/test_gpu_box_model.py -x -q -m gpu > gpurun_out/sanitizer_memcheck_new_pytest.log 2>&1
echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck_new.log; tail -n 3 gpurun_out/sanitizer_memcheck_new_pytest.log; grep "ERROR SUMMARY" gpurun_out/sanitizer_memcheck_new.log | sort | uniq -c | head
