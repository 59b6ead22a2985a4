#!/bin/bash
# One `ncu --set full` capture of a workload's top kernel + the timing decomposition of the tensor engine
# (full / CUDA-core side alone / MMA stream alone).  Usage: gpurun -- bash scripts/gpu_prof1.sh <tag> <workload> <kernel regex>
cd "$(dirname "$0")/.."
tag=${1:-prof1}; w=${2:-c2}; regex=${3:-tc_row_kernel}
out=gpurun_out/$tag; mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:$regex -s 3 -c 1 -o $out/prof_$w \
    python bench.py --steps 2 --warmup 3 --no-cpu --extra '' --workload $w --rk-steps 20 > $out/prof_$w.log 2>&1
for dbg in ${DBGS:-0 1 64}; do
  DDD1D_TC_DEBUG=$dbg timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --rk-steps 50 --no-cpu --extra '' > $out/${w}_d$dbg.json 2> $out/${w}_d$dbg.err
  python - <<PY
import json
try:
  d=json.load(open('$out/${w}_d$dbg.json')); print('$w debug=$dbg', 'ms/50 steps %.2f'%d['ms_per_step'], '%.3e gps/s'%d['value'])
except Exception as e: print('$w $dbg failed', e)
PY
done
