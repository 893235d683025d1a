"""Time the L mode (opts->L / opts->complete) on full-size BASELINE configs and check the factorization
(development helper; numbers quoted in DESIGN.md 5.6)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import util, oracle, spasm_b200
from spasm_b200 import synthetic, host
L = spasm_b200.lib()
L.spasm_b200_set_verbose(0)
mk = {'c1': lambda: synthetic.config1(), 'c2': lambda: synthetic.config2().transposed(), 'c4': lambda: synthetic.config4(1.0), 'c5': lambda: synthetic.config5()}
for w in sys.argv[1:] or ['c1']:
    t = mk[w]()
    A = host.compress(L, t)
    for mode in ('plain', 'L', 'complete'):
        for rep in range(2):
            o = host.default_opts(L)
            if mode == 'L':
                o.L = True
            if mode == 'complete':
                o.complete = True
            oracle.reset_rand()
            t0 = time.time(); f = host.echelonize(L, A, o); t1 = time.time()
        ok = all(host.factorization_verify(L, A, f, s) for s in (42, 1337)) if mode != 'plain' else None
        lnz = int(f.ptr.contents.L.contents.p[f.ptr.contents.L.contents.n]) if mode != 'plain' else 0
        print(f"{w} {t.n}x{t.m} mode={mode} rank {f.rank} wall {t1 - t0:.3f}s nnz(U) {int(f.ptr.contents.U.contents.p[f.rank])} nnz(L) {lnz} verify {ok}", flush=True)
