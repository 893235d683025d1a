import sys, os, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
os.environ['OMP_NUM_THREADS'] = '1'
import numpy as np
import util, oracle, spasm_b200
from spasm_b200 import synthetic
L = spasm_b200.lib()
L.spasm_b200_set_verbose(int(os.environ.get('V', '0')))
cases = [("tiny", synthetic.config1(0.002), {}), ("c1s", synthetic.config1(0.05), {}), ("c2s", synthetic.config2(0.02).transposed(), {}),
         ("c3s", synthetic.config3(0.01), dict(sparsity_threshold=0.01)), ("c4s", synthetic.config4(0.004), {}), ("c5s", synthetic.config5(0.03), {}),
         ("c1m", synthetic.config1(0.25), {})]
for name, t, kw in cases:
    t0 = time.time(); want = util.run_oracle(t, **kw); t1 = time.time()
    got = util.run_product(L, t, **kw); t2 = time.time()
    ok = {k: got[k] == want[k] for k in util.COMPARED + ("pairs_per_round",)}
    print(name, t.n, t.m, 'rank', got['rank'], want['rank'], 'found', got['found'], want['found'], 'finish', got['finish'], want['finish'], ok, f'oracle {t1-t0:.2f}s gpu {t2-t1:.2f}s', flush=True)
    if not ok['blocks']: print('  blocks', got['blocks'][:6], want['blocks'][:6])
