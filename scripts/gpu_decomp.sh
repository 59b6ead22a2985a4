#!/bin/bash
# Timing decomposition of the tensor engine on one box: full kernel, CUDA-core side alone (DDD1D_TC_DEBUG=1: no
# MMAs), MMA stream alone (=64: the teams sit the launch out), per operand precision.
# Usage: gpurun -- bash scripts/gpu_decomp.sh <tag> [workloads...]
tag=${1:-decomp}; shift
wls=${@:-c2}
out=gpurun_out/$tag
mkdir -p $out
for w in $wls; do
 for eng in ${ENGS:-tensor tensor_f16x2 tensor_f16}; do
  for dbg in ${DBGS:-0 1 64}; do
    DDD1D_TC_DEBUG=$dbg timeout 300 python bench.py --workload $w --engine $eng --steps 5 --warmup 3 --rk-steps 50 --no-cpu --extra '' > $out/${w}_${eng}_d$dbg.json 2> $out/${w}_${eng}_d$dbg.err
    python - <<PY
import json
try:
  d=json.load(open('$out/${w}_${eng}_d$dbg.json')); print('$w $eng debug=$dbg', 'ms/50 steps %.2f'%d['ms_per_step'], '%.3e gps/s'%d['value'])
except Exception as e: print('$w $eng $dbg failed', e)
PY
  done
 done
done
