#!/bin/bash
# r5 visit b: column ensembles as a replayed graph (BiogeochemicalModel.run(graph=True)) — sediment / sinking / NPD tests, timing of configs[1] as a run
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sediment.py tests/test_gpu_sinking.py tests/test_gpu_npd.py tests/test_gpu_kelp.py tests/test_gpu_gas_exchange.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 25
timeout 600 python scripts/time_column_ensemble.py 2>&1 | tail -n 3 | tee gpurun_out/time_column_ensemble.json
