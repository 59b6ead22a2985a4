#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_tensor.py -q --timeout 300 -p no:cacheprovider > gpurun_out/tc_tests.log 2>&1
echo "tc tests exit $?" >> gpurun_out/tc_tests.log
for w in c2 c3 c4; do
  DDD1D_ENGINE=tensor timeout -k 10 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/var_${w}_0.json 2> gpurun_out/var_${w}_0.err
done
echo done
