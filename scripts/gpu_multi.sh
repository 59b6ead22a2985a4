#!/bin/bash
# 2-GPU pass (gpurun --gpus 2): the sharded-evaluation test, then torchrun bench lines: weak, strong, reference arm.
cd "$(dirname "$0")/.."
out=gpurun_out/${1:-multi}; mkdir -p $out
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -q -x --timeout 300 -p no:cacheprovider > $out/pytest_multi.log 2>&1
echo "pytest exit $?" >> $out/pytest_multi.log; tail -2 $out/pytest_multi.log
run() { name=$1; port=$2; shift 2
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus 2 "$@" > $out/$name.json 2> $out/$name.err
  echo "$name exit $?" >> $out/$name.err
  python -c "
import json; d=json.loads(open('$out/$name.json').read().strip().splitlines()[-1]); print('$name', d.get('value'), (d.get('e2e') or {}).get('value'), d.get('scaling'), d.get('gather'))"
}
run bench_n2_weak 29511 --steps 20 --warmup 5 --no-cpu --extra ''
run bench_n2_strong 29512 --steps 20 --warmup 5 --scaling strong --no-cpu --extra ''
run bench_ref_n2 29513 --impl reference --steps 2 --warmup 1
