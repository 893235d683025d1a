#!/bin/bash
for v in 1024 4096 8192 16384 32768 65536; do echo "narrow_items=$v"; SPASM_B200_NARROW_ITEMS=$v REPS=2 timeout 300 python tools/gpu_full.py c2 2>&1 | grep -o "wall.*solve [0-9.]*ms" | tail -1; done
