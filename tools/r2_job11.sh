#!/bin/bash
# round 2, job 11 (8 GPUs): bench at N = 8 (scale leg), multi-rank parity at 8 ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/j11_bench_n8.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29642 tools/mgpu_check.py > gpurun_out/j11_mgpu8.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/j11_bench_n4.txt 2>&1
tail -1 gpurun_out/j11_bench_n8.txt | cut -c1-2500; tail -1 gpurun_out/j11_mgpu8.txt
