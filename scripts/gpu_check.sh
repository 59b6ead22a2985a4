#!/bin/bash
# GPU check of the tensor engine: parity tests in separate processes (a fault in one group must not poison
# the next), then short bench lines.  Usage: gpurun -- bash scripts/gpu_check.sh <tag>
tag=${1:-check}
out=gpurun_out/$tag
mkdir -p $out
run() { name=$1; shift; timeout 600 "$@" > $out/$name.log 2>&1; echo "$name rc=$?"; tail -3 $out/$name.log; }
run tensor_basic python -m pytest tests/test_gpu_tensor.py -x -q --timeout 300 -k "probe or per_call or trajectories and not packed"
run tensor_other python -m pytest tests/test_gpu_tensor.py -x -q --timeout 300 -k "other_nets or rejects or reduced"
run tensor_packed python -m pytest tests/test_gpu_tensor.py -x -q --timeout 300 -k "packed"
run long_horizon python -m pytest tests/test_gpu_long_horizon.py -x -q -s --timeout 400
for w in c2 c3 c4; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --rk-steps 100 --no-cpu --extra '' > $out/bench_$w.json 2> $out/bench_$w.err
  python - <<PY
import json
try:
  d=json.load(open('$out/bench_$w.json')); print('$w', '%.3e'%d['value'], 'ms', round(d['ms_per_step'],2), d['engine'], d.get('parity_check'))
except Exception as e: print('$w failed', e)
PY
done
