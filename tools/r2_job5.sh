#!/bin/bash
# round 2, job 9 (2 GPUs): multi-rank parity, bench at N = 2 (scale leg) and N = 1, the 2-rank GPU test, GPLU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/mgpu_check.py > gpurun_out/j9_mgpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/j9_bench_n2.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/j9_bench_n1.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_parity.py -q -k "multi or gplu or sparse or forced" 2>&1 | tail -8 > gpurun_out/j9_pytest.txt
tail -2 gpurun_out/j9_mgpu.txt; tail -1 gpurun_out/j9_bench_n2.txt | cut -c1-1800; tail -3 gpurun_out/j9_pytest.txt
