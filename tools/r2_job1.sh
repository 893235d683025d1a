#!/bin/bash
# round 2, job 1: windowed greedy search -- parity (shadow against the ordered kernel), timing, int8 peak
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( SPASM_B200_GREEDY_SHADOW=1 timeout 600 python tools/gpu_quick.py ) > gpurun_out/j1_quick.txt 2>&1
( SPASM_B200_GREEDY_SHADOW=1 REPS=1 C3SCALE=1.0 C4SCALE=1.0 timeout 900 python tools/gpu_full.py c2 c1 c3 c4 c5 ) > gpurun_out/j1_shadow_full.txt 2>&1
( REPS=3 C3SCALE=1.0 C4SCALE=1.0 SPASM_B200_TRACE=1 timeout 600 python tools/gpu_full.py c2 c1 c3 c4 ) > gpurun_out/j1_time_win.txt 2>&1
( REPS=3 SPASM_B200_GREEDY_WINDOW=1024 timeout 600 python tools/gpu_full.py c2 c1 ) > gpurun_out/j1_time_win1024.txt 2>&1
( REPS=3 SPASM_B200_GREEDY_WINDOW=256 timeout 600 python tools/gpu_full.py c2 c1 ) > gpurun_out/j1_time_win256.txt 2>&1
( REPS=3 C3SCALE=1.0 SPASM_B200_GREEDY_JOURNAL=1 timeout 600 python tools/gpu_full.py c2 c1 c3 ) > gpurun_out/j1_time_journal.txt 2>&1
timeout 300 python tools/int8_peak.py > gpurun_out/j1_int8.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/j1_launches_c2.csv python tools/gpu_full.py c2 > gpurun_out/j1_ncu.log 2>&1
tail -3 gpurun_out/j1_quick.txt gpurun_out/j1_shadow_full.txt gpurun_out/j1_time_win.txt
