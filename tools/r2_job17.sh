#!/bin/bash
# round 2, job 17 (1 GPU): per-launch durations of the panel kernels on config 1 (sampled factorisation)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=1
REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_rref -c 1400 --csv --log-file gpurun_out/j17_panel.csv python tools/gpu_full.py c1 > gpurun_out/j17.log 2>&1
python - <<'PY'
import csv, collections
rows=[l for l in open('gpurun_out/j17_panel.csv') if not l.startswith('==')]
agg=collections.defaultdict(list)
for r in csv.DictReader(rows):
    name=r['Kernel Name'].split('(')[0]
    v=float(r['Metric Value'].replace(',',''))
    u=r['Metric Unit']
    v*= {'ns':1e-3,'us':1,'ms':1e3}.get(u,1)
    agg[name].append(v)
for k,v in agg.items():
    v2=sorted(v)
    print(k, len(v), 'mean %.1f us'%(sum(v)/len(v)), 'median %.1f'%v2[len(v2)//2], 'min %.1f max %.1f'%(v2[0],v2[-1]), 'n<10us', sum(1 for x in v if x<10))
PY
