#!/bin/bash
# Everything the end of a round needs from ONE gpurun call: GPU tests, smoke, bench snapshots (config 2/1/3/5 + the
# reference arm), then the profiling recipe (tools/profile_run.sh).  tools/summarize_profiles.py turns gpurun_out/ into profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_r1d.err | tail -1 > gpurun_out/bench_r1d.json
python bench.py --workload config1 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1d_config1.json
python bench.py --workload config3 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_r1d_config3.json
python bench.py --workload config5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_r1d_config5.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r1d_reference.json
LPS=$(python -c "import json; d=json.load(open('gpurun_out/bench_r1d.json')); print(d['gpu_launches']//d['steps'])")
L1=$(python -c "import json; d=json.load(open('gpurun_out/bench_r1d_config1.json')); print(d['gpu_launches']//d['steps'])")
LPS=$LPS L1=$L1 bash tools/profile_run.sh > gpurun_out/profile_run.log 2>&1
