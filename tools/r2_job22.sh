#!/bin/bash
# round 2, job 22 (1 GPU): warp-level M^-1 tail of the panel kernels: parity subset + timing
cd "$(dirname "$0")/.."
export OMP_NUM_THREADS=1
( REPS=3 timeout 60 python tools/gpu_full.py c2 c1 ) 2>&1 | grep -a "^c[0-9]" | cut -c1-175 > gpurun_out/j22_time.txt
timeout 100 python -m pytest tests/test_gpu_components.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/j22_pytest.txt
cat gpurun_out/j22_time.txt; tail -4 gpurun_out/j22_pytest.txt
