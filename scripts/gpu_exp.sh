#!/bin/bash
# timing experiments on the tensor engine (env switches), C2 launch shape
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout -k 10 200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err; python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/exp_$name.json').read().strip().splitlines()[-1]); print('$name', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'])
except Exception as e: print('$name failed', e)
PY
}
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  run $name $(echo $envs | tr ',' ' ')
done
