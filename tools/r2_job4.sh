#!/bin/bash
# round 2, job 4: greedy windows v4 (narrow-frontier walk, E rows in registers), hop statistics of the dataflow solve,
# one-shot tool timing, profiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( SPASM_B200_GREEDY_SHADOW=1 timeout 600 python tools/gpu_quick.py ) > gpurun_out/j4_quick.txt 2>&1
( SPASM_B200_GREEDY_SHADOW=1 REPS=1 C3SCALE=1.0 C4SCALE=1.0 timeout 900 python tools/gpu_full.py c2 c1 c3 c4 c5 ) > gpurun_out/j4_shadow_full.txt 2>&1
( REPS=3 C3SCALE=1.0 C4SCALE=1.0 timeout 600 python tools/gpu_full.py c2 c1 c3 c4 ) > gpurun_out/j4_time.txt 2>&1
( REPS=3 SPASM_B200_GREEDY_WINDOW=1024 timeout 600 python tools/gpu_full.py c2 c1 ) > gpurun_out/j4_time_win1024.txt 2>&1
( REPS=2 SPASM_B200_TRACE=1 timeout 600 python tools/gpu_full.py c2 c1 ) 2>&1 | grep -a "dataflow solve\|^c[0-9] " > gpurun_out/j4_hops.txt
timeout 600 python tools/oneshot.py b200_rank ref_rank > gpurun_out/j4_oneshot.txt 2>&1
timeout 1500 bash tools/r2_profile.sh > gpurun_out/j4_profile.log 2>&1
timeout 600 python -m pytest tests/test_gpu_sanitizer.py -q 2>&1 | tail -5 > gpurun_out/j4_sanitizer.txt
grep -a "^c[0-9] " gpurun_out/j4_time.txt | cut -c1-200
