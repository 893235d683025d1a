#!/bin/bash
# round 2, job 7 (1 GPU): chase + 32-byte records in the greedy search, stream priorities for the dense look-ahead; sanitizers; full GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( SPASM_B200_GREEDY_SHADOW=1 timeout 600 python tools/gpu_quick.py ) > gpurun_out/j7_quick.txt 2>&1
( SPASM_B200_GREEDY_SHADOW=1 REPS=1 C3SCALE=1.0 C4SCALE=1.0 timeout 900 python tools/gpu_full.py c2 c1 c4 c5 ) > gpurun_out/j7_shadow_full.txt 2>&1
( REPS=3 C3SCALE=1.0 C4SCALE=1.0 timeout 600 python tools/gpu_full.py c2 c1 c3 c4 c5 ) > gpurun_out/j7_time.txt 2>&1
( REPS=3 SPASM_B200_GREEDY_WINDOW=1024 timeout 600 python tools/gpu_full.py c2 ) > gpurun_out/j7_time_win1024.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/j7_launches_c2.csv python tools/gpu_full.py c2 > gpurun_out/j7_ncu.log 2>&1
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/j7_pytest.txt
grep -a "^c[0-9] " gpurun_out/j7_time.txt | cut -c1-200; tail -5 gpurun_out/j7_pytest.txt
