#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; w=$2; shift; shift; env "$@" timeout -k 10 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err; python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/exp_$name.json').read().strip().splitlines()[-1]); print('$name', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'])
except Exception as e: print('$name failed', e)
PY
}
for w in c2 c3 c4; do
  for i in 1 2 3 4; do
    run ${w}_i$i $w DDD1D_TC_ISSUERS=$i
    run ${w}_i${i}_sleep $w DDD1D_TC_ISSUERS=$i DDD1D_TC_DEBUG=4
  done
done
