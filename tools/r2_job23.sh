#!/bin/bash
# round 2, job 23 (1 GPU, last seconds of the budget): the 1024-thread sampled panel kernel (opt-in) -- dense parity + timing
cd "$(dirname "$0")/.."
export OMP_NUM_THREADS=1 SPASM_B200_PANEL_WIDE=1
( REPS=3 timeout 40 python tools/gpu_full.py c2 c1 ) 2>&1 | grep -a "^c[0-9]" | cut -c1-175 > gpurun_out/j23_time.txt
timeout 40 python -m pytest tests/test_gpu_components.py -m gpu -q -x -k "dense_rref or gemm" 2>&1 | tail -4 > gpurun_out/j23_pytest.txt
cat gpurun_out/j23_time.txt; tail -3 gpurun_out/j23_pytest.txt
