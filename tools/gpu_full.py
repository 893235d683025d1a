"""Run full-size configs through the product and print the phase timers (development helper)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import util, oracle, spasm_b200
from spasm_b200 import synthetic, host
L = spasm_b200.lib()
L.spasm_b200_set_verbose(int(os.environ.get('V', '0')))
which = sys.argv[1:] or ['c2']
mk = {'c1': lambda: (synthetic.config1(), {}), 'c2': lambda: (synthetic.config2().transposed(), {}),
      'c3': lambda: (synthetic.config3(float(os.environ.get('C3SCALE', '0.25'))), dict(sparsity_threshold=0.01)),
      'c4': lambda: (synthetic.config4(float(os.environ.get('C4SCALE', '0.2'))), {}), 'c5': lambda: (synthetic.config5(), {})}
for w in which:
    t, kw = mk[w]()
    A = host.compress(L, t)
    for rep in range(int(os.environ.get('REPS', '2'))):
        oracle.reset_rand(); L.spasm_b200_reset_stats()
        t0 = time.time(); f = host.echelonize(L, A, host.default_opts(L, **kw)); t1 = time.time()
        s = util.product_stats(L)
        print(w, t.n, t.m, 'rank', f.rank, f'wall {t1-t0:.3f}s', 'found', [(s.found_FL[r], s.found_FLcol[r], s.found_greedy[r]) for r in range(s.nrounds)],
              f'pivots {s.ms_pivots:.1f}ms (greedy {s.ms_pivots_greedy:.1f}) solve {s.ms_solve:.1f}ms dense {s.ms_dense:.1f}ms (gemm {s.ms_dense_gemm:.1f}) launches {s.kernel_launches} depth {s.dag_depth} edges {s.greedy_edges}',
              'blocks', [(s.block_Sn[k], s.block_Sm[k], s.block_rr[k], s.block_w[k]) for k in range(min(s.nblocks, 8))], flush=True)
    if os.environ.get('RREF'):
        t0 = time.time(); Rm, _ = host.rref(L, f); t1 = time.time(); Km = host.kernel(L, f); t2 = time.time()
        print('  rref', f'{t1-t0:.3f}s nnz {Rm.nnz}', 'kernel', f'{t2-t1:.3f}s rows {Km.n}', flush=True)
