#!/bin/bash
# round 2, job 6 (1 GPU): dataflow solve with the dependents' blob, dense look-ahead -- parity + timing; racecheck; full GPU test suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( SPASM_B200_GREEDY_SHADOW=1 timeout 600 python tools/gpu_quick.py ) > gpurun_out/j6_quick.txt 2>&1
( REPS=3 C3SCALE=1.0 C4SCALE=1.0 timeout 600 python tools/gpu_full.py c2 c1 c3 c4 c5 ) > gpurun_out/j6_time.txt 2>&1
( REPS=3 C3SCALE=1.0 C4SCALE=1.0 SPASM_B200_NO_LOOKAHEAD=1 SPASM_B200_FLOW1=1 timeout 600 python tools/gpu_full.py c2 c1 c3 c4 c5 ) > gpurun_out/j6_time_old.txt 2>&1
( REPS=2 SPASM_B200_TRACE=1 timeout 600 python tools/gpu_full.py c2 c1 ) 2>&1 | grep -a "dataflow solve\|^c[0-9] " > gpurun_out/j6_hops.txt
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/j6_pytest.txt
grep -a "^c[0-9] " gpurun_out/j6_time.txt | cut -c1-200; tail -3 gpurun_out/j6_pytest.txt
