"""Microbenchmark of the modular product C -= A*B (tools only; run on the GPU box).
usage: python tools/gemm_bench.py [prime]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spasm_b200

lib = spasm_b200.lib()
prime = int(sys.argv[1]) if len(sys.argv) > 1 else 42013
shapes = [(1000, 1000, 6813), (1024, 6813, 1000), (1000, 23800, 1000), (1000, 23800, 8000), (4096, 4096, 4096), (8192, 8192, 8192)]
names = {0: "cuda-core", 1: "tensor(pack A+B)", 2: "tensor(pack A)", 3: "tensor kernel"}
for (M, N, K) in shapes:
    line = "%5d x %5d x %5d p=%d:" % (M, N, K, prime)
    for mode in (0, 1, 2, 3):
        if mode == 0 and M * N * K > 3e11:
            continue
        ms = lib.spasm_b200_gemm_time(prime, M, N, K, mode, 5)
        line += "  %s %.3f ms (%.1f Tfieldop/s)" % (names[mode], ms, 2.0 * M * N * K / ms / 1e9)
    print(line, flush=True)
