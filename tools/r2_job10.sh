#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( REPS=2 C4SCALE=1.0 RREF=1 SPASM_B200_TRACE=1 timeout 900 python tools/gpu_full.py c4 ) 2>&1 | grep -a "rref\|kernel\|^c4\|trace\] core" > gpurun_out/j10_rref_trace.txt
( REPS=1 C4SCALE=1.0 RREF=1 SPASM_B200_TRACE=1 SPASM_B200_PANEL_GB=16 timeout 900 python tools/gpu_full.py c4 ) 2>&1 | grep -a "rref\|kernel\|^c4" > gpurun_out/j10_rref_trace_16gb.txt
tail -30 gpurun_out/j10_rref_trace.txt
