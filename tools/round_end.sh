#!/bin/bash
# Everything the end of a round needs from ONE gpurun call on one GPU: GPU tests, smoke, bench snapshots (the default line,
# one line per BASELINE config, the reference arm), then the profiling recipe (tools/r2_profile.sh).
# tools/summarize_profiles.py turns gpurun_out/ into profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r2}
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_$TAG.json
python bench.py --workload all --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_${TAG}_all.jsonl
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_${TAG}_reference.json
timeout 1500 bash tools/r2_profile.sh > gpurun_out/profile_run.log 2>&1
ls -la gpurun_out/
