#!/bin/bash
cd "$(dirname "$0")/.."
export OMP_NUM_THREADS=1
REPS=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rref|k_gemm_sub|k_copy_pivot" -c 1200 --csv --log-file gpurun_out/j21_panel.csv python tools/gpu_full.py c2 > gpurun_out/j21.log 2>&1
python - <<'PY'
import csv, collections
rows=[l for l in open('gpurun_out/j21_panel.csv') if not l.startswith('==')]
agg=collections.defaultdict(list)
for r in csv.DictReader(rows):
    name=r['Kernel Name'].split('(')[0]+' grid'+r['Grid Size']
    v=float(r['Metric Value'].replace(',',''))
    v*= {'ns':1e-3,'us':1,'ms':1e3}.get(r['Metric Unit'],1)
    agg[name].append(v)
for k,v in agg.items():
    v2=sorted(v)
    print(k, len(v), 'mean %.1f us'%(sum(v)/len(v)), 'median %.1f'%v2[len(v2)//2], 'min %.1f max %.1f'%(v2[0],v2[-1]), 'sum %.2f ms'%(sum(v)/1e3))
PY
