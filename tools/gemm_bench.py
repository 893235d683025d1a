"""Microbenchmark of the modular product C -= A*B (tools only; run on the GPU box).
usage: python tools/gemm_bench.py [prime [M N K [mode]]]     (mode: see spasm_b200_gemm_time in include/spasm_b200.h)"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spasm_b200

lib = spasm_b200.lib()
prime = int(sys.argv[1]) if len(sys.argv) > 1 else 42013
modes = (0, 1, 2, 3)
if len(sys.argv) > 4:
    only = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    modes = (int(sys.argv[5]),) if len(sys.argv) > 5 else modes
shapes = [only] if len(sys.argv) > 4 else [(1000, 1000, 6813), (1024, 6813, 1000), (1000, 23800, 1000), (1000, 23800, 8000), (4096, 4096, 4096), (8192, 8192, 8192)]
names = {0: "dense_gemm_sub (dispatch)", 1: "tensor(pack A+B)", 2: "tensor(pack A)", 3: "tensor kernel"}
for (M, N, K) in shapes:
    line = "%5d x %5d x %5d p=%d:" % (M, N, K, prime)
    for mode in modes:
        if mode == 0 and M * N * K > 3e11:
            continue
        ms = lib.spasm_b200_gemm_time(prime, M, N, K, mode, 5)
        line += "  %s %.3f ms (%.1f Tfieldop/s)" % (names[mode], ms, 2.0 * M * N * K / ms / 1e9)
    print(line, flush=True)
