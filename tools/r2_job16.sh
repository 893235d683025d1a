#!/bin/bash
# round 2, job 16 (1 GPU): sampled panel factorisation (parity + A/B timing), L mode with the first-round shortcut
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
timeout 900 python -m pytest tests/test_gpu_components.py tests/test_gpu_lu.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -25 > gpurun_out/j16_pytest.txt
( REPS=3 C3SCALE=1.0 C4SCALE=1.0 timeout 600 python tools/gpu_full.py c2 c1 c3 c4 c5 ) 2>&1 | grep -a "^c[0-9]" | cut -c1-200 > gpurun_out/j16_time.txt
( REPS=3 SPASM_B200_PANEL_NO_SAMPLE=1 timeout 300 python tools/gpu_full.py c1 ) 2>&1 | grep -a "^c[0-9]" | cut -c1-200 > gpurun_out/j16_time_nosample.txt
( timeout 300 python tools/lu_time.py c4 ) 2>/dev/null | grep -a "mode=" > gpurun_out/j16_lu_time.txt
tail -6 gpurun_out/j16_pytest.txt; cat gpurun_out/j16_time.txt gpurun_out/j16_time_nosample.txt gpurun_out/j16_lu_time.txt
