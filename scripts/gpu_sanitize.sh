#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (small grids only): memcheck on everything but the full-size
# file, racecheck on the kernels that use shared memory (PAR scans, negative scaling / Ω prologue).
set -u
mkdir -p gpurun_out
FILES=$(ls tests/test_gpu_*.py | grep -v full_size | grep -v box_model)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck.log python -m pytest $FILES -x -q -m gpu > gpurun_out/sanitizer_memcheck_pytest.log 2>&1
echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck.log; tail -3 gpurun_out/sanitizer_memcheck_pytest.log; grep -c "ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log; grep "ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log | sort | uniq -c | head
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitizer_racecheck.log python -m pytest tests/test_gpu_light.py tests/test_gpu_negs.py tests/test_gpu_pisces.py -x -q -m gpu > gpurun_out/sanitizer_racecheck_pytest.log 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/sanitizer_racecheck.log; tail -3 gpurun_out/sanitizer_racecheck_pytest.log; grep "RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck.log | sort | uniq -c | head
