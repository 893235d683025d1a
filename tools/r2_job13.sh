#!/bin/bash
# round 2, job 13 (1 GPU): GPU ingest, bulk-copy (TMA) staging in the dataflow solve (parity + A/B timing), rank --certificate,
# the two L-path programs the previous job did not reach
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
( timeout 300 python tools/gpu_quick.py ) > gpurun_out/j13_quick.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_ingest.py -m gpu -q -s 2>&1 | grep -a "ingest\]\|passed\|failed\|Error\|error" | tail -20 > gpurun_out/j13_pytest_ingest.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_components.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/j13_pytest_parity.txt
timeout 300 python -m pytest tests/test_reference_tests.py -m gpu -q -k "rank_cert or dense_lu_ffpack" 2>&1 | tail -8 > gpurun_out/j13_pytest_ref.txt
( REPS=4 timeout 300 python tools/gpu_full.py c2 c1 c4 ) 2>&1 | grep -a "^c[0-9]" | cut -c1-260 > gpurun_out/j13_time_bulk.txt
( REPS=4 SPASM_B200_FLOW2_NO_BULK=1 timeout 300 python tools/gpu_full.py c2 c1 c4 ) 2>&1 | grep -a "^c[0-9]" | cut -c1-260 > gpurun_out/j13_time_nobulk.txt
tail -3 gpurun_out/j13_quick.txt; cat gpurun_out/j13_pytest_ingest.txt; tail -4 gpurun_out/j13_pytest_parity.txt; tail -3 gpurun_out/j13_pytest_ref.txt
cut -c1-170 gpurun_out/j13_time_bulk.txt; cut -c1-170 gpurun_out/j13_time_nobulk.txt
