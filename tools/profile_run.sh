#!/bin/bash
# Profiling recipe of this repository (B200_PROFILING.md), run on a B200 through gpurun:
#   1. launch list (gpu__time_duration) of one steady-state step of the bench workload (config 2) and of config 1
#   2. `ncu --set full` captures of the dominant kernels
# Outputs land in gpurun_out/ (scratch); tools/summarize_profiles.py turns them into profiles/*.md.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LPS=${LPS:-381}     # launches per echelonize step of config 2 (bench prints gpu_launches / steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * LPS)) -c $LPS --csv --log-file gpurun_out/launches_config2.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * ${L1:-1218})) -c ${L1:-1218} --csv --log-file gpurun_out/launches_config1.csv \
    python bench.py --workload config1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench1_under_ncu.log 2>&1
for k in k_greedy_ooo k_panel_solve_flow k_kahn_async k_rref_panel_cluster; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$k.log 2>&1
done
# the tensor-core product: config 1 has the 1000 x 6813 x 1000 block reductions (128 x 64 tiles, 2 CTAs per SM) ...
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_gemm_packed -s 20 -c 1 -f -o gpurun_out/prof_k_umma_gemm_packed \
    python bench.py --workload config1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_umma.log 2>&1
# ... and a large product (8192^3, 2 limbs, 128 x 128 tiles) for the tensor-pipe ceiling of the kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_gemm_packed -s 2 -c 1 -f -o gpurun_out/prof_k_umma_gemm_packed_8192 \
    python tools/gemm_bench.py 42013 8192 8192 8192 3 > gpurun_out/prof_umma_8192.log 2>&1
timeout 300 python tools/gemm_bench.py 42013 > gpurun_out/gemm_bench_42013.txt 2>&1
timeout 300 python tools/gemm_bench.py 65537 > gpurun_out/gemm_bench_65537.txt 2>&1
timeout 300 python tools/gemm_bench.py 2147483629 > gpurun_out/gemm_bench_2147483629.txt 2>&1
ls -la gpurun_out/
