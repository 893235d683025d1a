#!/bin/bash
# round 2, job 12 (1 GPU): the L path (opts->L / complete, spasm_ffpack_LU, solve / gesv / certificates)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
timeout 900 python -m pytest tests/test_gpu_lu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/j12_pytest_lu.txt
timeout 1200 python -m pytest tests/test_reference_tests.py -m gpu -q -k "lu or solve or gesv or rank_cert" 2>&1 | tail -30 > gpurun_out/j12_pytest_ref.txt
tail -15 gpurun_out/j12_pytest_lu.txt; tail -8 gpurun_out/j12_pytest_ref.txt
