import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import util, oracle, spasm_b200
from spasm_b200 import synthetic
L = spasm_b200.lib(); L.spasm_b200_set_verbose(1)
t = synthetic.uniform_rows(1714, 473, 11, seed=20240302, values="small", distinct=False)
want = util.run_oracle(t, sparsity_threshold=0.01)
print('oracle', want['rank'], want['found'], flush=True)
got = util.run_product(L, t, sparsity_threshold=0.01)
print('gpu', got['rank'], got['found'], {k: got[k] == want[k] for k in util.COMPARED + ('pairs_per_round',)})
