#!/bin/bash
# round 2, job 18 (1 GPU): final state -- the whole GPU suite, bench lines, then per-launch durations of the panel kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/j18_pytest.txt 2>&1
unset OMP_NUM_THREADS
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | grep -a "^{" > gpurun_out/j18_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep -a "^{" > gpurun_out/j18_bench_reference.json
timeout 600 python bench.py --workload all --no-cpu-baseline --no-scale-leg --no-schur-leg --steps 5 --warmup 3 2>/dev/null | grep -a "^{" > gpurun_out/j18_bench_all.json
export OMP_NUM_THREADS=1
bash tools/r2_job17.sh > gpurun_out/j18_panel.txt 2>&1
tail -8 gpurun_out/j18_pytest.txt; cut -c1-300 gpurun_out/j18_bench.json; tail -6 gpurun_out/j18_panel.txt
