#!/bin/bash
# round 2, job 8 (1 GPU): timing after the chase was switched off, dataflow solve with prefetches and two groups per thread,
# the sparse GPLU finisher, full-size goldens
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( SPASM_B200_GREEDY_SHADOW=1 timeout 600 python tools/gpu_quick.py ) > gpurun_out/j8_quick.txt 2>&1
( REPS=3 C3SCALE=1.0 C4SCALE=1.0 timeout 600 python tools/gpu_full.py c2 c1 c3 c4 c5 ) > gpurun_out/j8_time.txt 2>&1
( REPS=3 SPASM_B200_GREEDY_JOURNAL=1 timeout 600 python tools/gpu_full.py c2 c1 c4 ) > gpurun_out/j8_time_journal.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sanitizer.py -q -k "gplu or sparse or full_size or sanitizer or forced" 2>&1 | tail -8 > gpurun_out/j8_pytest.txt
grep -a "^c[0-9] " gpurun_out/j8_time.txt | cut -c1-200; tail -4 gpurun_out/j8_pytest.txt
