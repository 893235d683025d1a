#!/bin/bash
# round 2, job 14 (2 GPUs): multi-rank parity incl. the gather-to-root mode, bench at N = 2 (scale leg in both modes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 tools/mgpu_check.py 2>&1 | grep -av "Compressing\|actual NZ" | tail -60 > gpurun_out/j14_mgpu2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | grep -av "Compressing\|actual NZ" | tail -40 > gpurun_out/j14_bench_n2.txt
tail -8 gpurun_out/j14_mgpu2.txt; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/j14_bench_n2.txt') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','nccl_bytes_per_step')}, d['e2e'])
print(json.dumps(d.get('scale_leg'), indent=1))
PY
