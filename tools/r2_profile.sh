#!/bin/bash
# Round-2 profiling recipe (B200_PROFILING.md), one gpurun call on one GPU:
#   1. launch lists WITH dram bytes of one steady-state echelonize of config 2 / config 1 / config 3
#      (gpu__time_duration + dram__bytes_read/write per launch: the per-step traffic of every kernel family)
#   2. `ncu --set full` captures of the dominant kernels (a late window of the greedy search, the dataflow solve, the panel)
# Outputs land in gpurun_out/ (scratch); tools/summarize_profiles.py turns them into profiles/r2_*.md + traffic.json.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for w in c2 c1 c3; do
  REPS=2 C3SCALE=1.0 timeout 900 ncu --metrics $M --clock-control none -c 20000 --csv --log-file gpurun_out/r2_launches_$w.csv \
      python tools/gpu_full.py $w > gpurun_out/r2_launches_$w.log 2>&1
done
# full captures: greedy search windows in the middle of the run (launch 40 of each kernel), the solve, the dense panel
for k in k_win_bfs k_win_resolve; do
  REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o gpurun_out/r2_prof_$k \
      python tools/gpu_full.py c2 > gpurun_out/r2_prof_$k.log 2>&1
done
for k in k_panel_solve_flow2 k_rref_panel_cluster; do
  REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 -f -o gpurun_out/r2_prof_$k \
      python tools/gpu_full.py c2 > gpurun_out/r2_prof_$k.log 2>&1
done
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_gemm_packed -s 20 -c 1 -f -o gpurun_out/r2_prof_k_umma_gemm_packed \
    python tools/gpu_full.py c1 > gpurun_out/r2_prof_umma.log 2>&1
ls -la gpurun_out/ | grep r2_
