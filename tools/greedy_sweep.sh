#!/bin/bash
# development helper: sweep the greedy kernel's launch shape on config 2
for cfg in "256 1 8" "128 1 8" "64 1 8" "256 0 8" "128 0 16" "64 0 32" "32 0 32" "32 0 64"; do
  set -- $cfg
  echo "threads=$1 smem=$2 per_sm=$3"
  SPASM_B200_GREEDY_THREADS=$1 SPASM_B200_GREEDY_SMEM=$2 SPASM_B200_GREEDY_PER_SM=$3 REPS=1 timeout 300 python tools/gpu_full.py c2 2>&1 | grep -o "found.*edges [0-9]*"
done
