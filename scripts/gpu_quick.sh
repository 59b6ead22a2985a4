#!/bin/bash
# parity tests + C2/C3/C4 bench lines (no CPU leg): the loop used while tuning the tensor engine
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
bash scripts/gpu_exp.sh c2:W=c2 ${@}
for w in c3 c4; do
  timeout -k 10 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'])"
done
