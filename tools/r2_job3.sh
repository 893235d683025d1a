#!/bin/bash
# round 2, job 3: windowed greedy v3 + dataflow solve v2 -- parity, timing, reference tests, sanitizers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( SPASM_B200_GREEDY_SHADOW=1 timeout 600 python tools/gpu_quick.py ) > gpurun_out/j3_quick.txt 2>&1
( SPASM_B200_GREEDY_SHADOW=1 REPS=1 C3SCALE=1.0 C4SCALE=1.0 timeout 900 python tools/gpu_full.py c2 c1 c3 c4 c5 ) > gpurun_out/j3_shadow_full.txt 2>&1
( REPS=3 C3SCALE=1.0 C4SCALE=1.0 timeout 600 python tools/gpu_full.py c2 c1 c3 c4 ) > gpurun_out/j3_time.txt 2>&1
( REPS=3 SPASM_B200_GREEDY_WINDOW=1024 timeout 600 python tools/gpu_full.py c2 c1 ) > gpurun_out/j3_time_win1024.txt 2>&1
( REPS=3 SPASM_B200_GREEDY_WINDOW=256 timeout 600 python tools/gpu_full.py c2 c1 ) > gpurun_out/j3_time_win256.txt 2>&1
( REPS=3 C4SCALE=1.0 SPASM_B200_FLOW1=1 timeout 600 python tools/gpu_full.py c2 c1 c4 ) > gpurun_out/j3_time_flow1.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/j3_launches_c2.csv python tools/gpu_full.py c2 > gpurun_out/j3_ncu.log 2>&1
timeout 1200 python -m pytest tests/test_reference_tests.py tests/test_gpu_fullsize.py -q 2>&1 | tail -15 > gpurun_out/j3_reftests.txt
timeout 1500 python -m pytest tests/test_gpu_sanitizer.py -q 2>&1 | tail -15 > gpurun_out/j3_sanitizer.txt
grep -a "^c[0-9] " gpurun_out/j3_time.txt | cut -c1-200
