#!/bin/bash
for v in 128 256 512 1024; do echo "flow_threads=$v"; SPASM_B200_FLOW_THREADS=$v REPS=2 timeout 300 python tools/gpu_full.py c2 2>&1 | grep -o "wall.*solve [0-9.]*ms" | tail -1; done
echo default; REPS=2 timeout 300 python tools/gpu_full.py c2 c1 2>&1 | grep -o "c[0-9] .*wall.*solve [0-9.]*ms" | tail -3
