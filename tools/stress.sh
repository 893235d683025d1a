#!/bin/bash
# development helper: repeat a full-size run in fresh processes and count failures
# usage: tools/stress.sh <config> <runs> [env assignments...]
cfg=$1; runs=$2; shift 2
ok=0; bad=0
for i in $(seq 1 $runs); do
  if env "$@" REPS=${STRESS_REPS:-3} timeout 600 python tools/gpu_full.py $cfg > gpurun_out/stress_$i.txt 2>&1; then ok=$((ok+1)); else bad=$((bad+1)); tail -3 gpurun_out/stress_$i.txt | cut -c1-300; fi
done
echo "stress $cfg $@: ok=$ok bad=$bad"
