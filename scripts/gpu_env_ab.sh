#!/bin/bash
# A/B of run-time switches of the current library: bench lines of WLS for each "name:ENV=val,ENV=val" spec, two rounds.
# Usage: gpurun -- bash scripts/gpu_env_ab.sh <tag> spec...
cd "$(dirname "$0")/.."
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
if [ -n "$TESTS" ]; then
  timeout -k 10 600 python -m pytest $TESTS -q -x --timeout 300 -p no:cacheprovider > $out/pytest.log 2>&1; echo "pytest: $(tail -1 $out/pytest.log)"
fi
for rep in 1 2; do
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  for w in ${WLS:-c2 c3 c4 c2s}; do
    env $(echo $envs | tr ',' ' ') timeout -k 10 200 python bench.py --workload $w --steps 5 --warmup 3 --rk-steps ${RK:-50} --no-cpu --extra "" > $out/${name}_${w}_$rep.json 2> $out/${name}_${w}_$rep.err
    python -c "
import json; d=json.loads(open('$out/${name}_${w}_$rep.json').read().strip().splitlines()[-1]); print('$name', '$w', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'], d.get('parity_check',{}).get('rel_err'))"
  done
done
done
