#!/bin/bash
# round 2, job 15 (1 GPU): the whole GPU suite, bench lines (default, reference arm, all workloads), L-mode timings,
# refreshed launch list of config 2 and full capture of the dataflow solve (bulk-copy staging)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( time timeout 2400 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -45 ) > gpurun_out/j15_pytest.txt 2>&1
unset OMP_NUM_THREADS
timeout 900 python bench.py --steps 10 --warmup 3 2>/dev/null | grep -a "^{" > gpurun_out/j15_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep -a "^{" > gpurun_out/j15_bench_reference.json
timeout 900 python bench.py --workload all --no-cpu-baseline --no-scale-leg --no-schur-leg --steps 5 --warmup 3 2>/dev/null | grep -a "^{" > gpurun_out/j15_bench_all.json
export OMP_NUM_THREADS=1
( timeout 600 python tools/lu_time.py c1 c4 ) 2>/dev/null | grep -a "mode=" > gpurun_out/j15_lu_time.txt
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
REPS=2 timeout 600 ncu --metrics $M --clock-control none -c 20000 --csv --log-file gpurun_out/r2_launches_c2.csv python tools/gpu_full.py c2 > gpurun_out/r2_launches_c2.log 2>&1
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_panel_solve_flow2 -s 0 -c 1 -f -o gpurun_out/r2_prof_k_panel_solve_flow2 python tools/gpu_full.py c2 > gpurun_out/r2_prof_k_panel_solve_flow2.log 2>&1
tail -12 gpurun_out/j15_pytest.txt; cut -c1-600 gpurun_out/j15_bench.json; cat gpurun_out/j15_lu_time.txt
