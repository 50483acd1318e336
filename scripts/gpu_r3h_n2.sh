#!/bin/bash
# r3 visit h (2 GPUs): the 2-rank NCCL test (slab stage + inventory all-reduce), the default bench line as the driver launches it at N = 2
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_n2.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4_n2.json 2> gpurun_out/bench_pisces_c4_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/bench_pisces_c4_n2.json | cut -c1-3000; tail -5 gpurun_out/bench_pisces_c4_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; cat gpurun_out/bench_ref_n2.json | cut -c1-1500; tail -3 gpurun_out/bench_ref_n2.err
