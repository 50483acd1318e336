#!/bin/bash
# r3 visit i (8 GPUs): the default bench line as the driver launches it at N = 8 (strong scaling of the one C4 grid, inventory all-reduce over NCCL,
# e2e with every rank copying at once, weak sub-record), then N = 4 without the e2e leg
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1; lscpu | head -30 > gpurun_out/lscpu_n8.txt; free -g > gpurun_out/free_n8.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_pisces_c4_n8.json 2> gpurun_out/bench_pisces_c4_n8.err; echo "bench n8 rc=$?"; cut -c1-2500 gpurun_out/bench_pisces_c4_n8.json; tail -5 gpurun_out/bench_pisces_c4_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_pisces_c4_n4.json 2> gpurun_out/bench_pisces_c4_n4.err; echo "bench n4 rc=$?"; cut -c1-600 gpurun_out/bench_pisces_c4_n4.json; tail -3 gpurun_out/bench_pisces_c4_n4.err
