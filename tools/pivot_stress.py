"""Development helper: repeat the structural pivot search on one matrix and report runs that differ from the first
(with SPASM_B200_GREEDY_SHADOW=1 the library also compares the out-of-order greedy kernel with the ordered one)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import util, oracle, spasm_b200
from spasm_b200 import synthetic, host
L = spasm_b200.lib()
L.spasm_b200_set_verbose(0)
which = sys.argv[1] if len(sys.argv) > 1 else 'c3'
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 50
t = {'c3': lambda: synthetic.config3(float(os.environ.get('C3SCALE', '1.0'))), 'c2': lambda: synthetic.config2().transposed(),
     'c1': lambda: synthetic.config1()}[which]()
A = host.compress(L, t)
first = None
for run in range(runs):
    npiv, p, fact = host.pivots_extract_structural(L, A)
    key = (npiv, p[:npiv].tobytes())
    if first is None:
        first = key
    print(run, npiv, 'same' if key == first else 'DIFFERENT', flush=True)
