#!/bin/bash
# Profiling recipe of this repository (B200_PROFILING.md): launch list of one steady-state bench step + full captures
# of the two dominant kernels.  Outputs land in gpurun_out/ (copy summaries to profiles/).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# 1. every launch of one steady-state step with its device time (3 warm-up steps = ~1800 launches are skipped)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1776} -c ${COUNT:-700} --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# 2. the two dominant kernels, once each, full metric set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_greedy -s 3 -c 1 -f -o gpurun_out/prof_greedy \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_greedy.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_panel_solve -s 13 -c 1 -f -o gpurun_out/prof_solve \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_solve.log 2>&1
ls -la gpurun_out/
