"""Development helper: a few medium inputs, several repetitions each (so that the memory pool recycles blocks),
meant to run under `compute-sanitizer --tool initcheck|memcheck|racecheck`."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import util, oracle, spasm_b200
from spasm_b200 import synthetic, host
L = spasm_b200.lib()
L.spasm_b200_set_verbose(0)
scale = float(os.environ.get('SCALE', '1'))
cases = [("c3", synthetic.config3(0.04 * scale), dict(sparsity_threshold=0.01)), ("c2", synthetic.config2(0.03 * scale).transposed(), {}),
         ("c1", synthetic.config1(0.1 * scale), {})]
for name, t, kw in cases:
    A = host.compress(L, t)
    for rep in range(int(os.environ.get('REPS', '2'))):
        oracle.reset_rand()
        f = host.echelonize(L, A, host.default_opts(L, **kw))
        print(name, rep, 'rank', f.rank, flush=True)
