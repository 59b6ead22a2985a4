#!/bin/bash
# tensor engine check: tests, then benches (two rows in flight per team vs one)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_tensor.py -q --timeout 300 -p no:cacheprovider > gpurun_out/tc_tests.log 2>&1
echo "tc tests exit $?" >> gpurun_out/tc_tests.log
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "parity exit $?" >> gpurun_out/pytest_gpu.log
DDD1D_ENGINE=tensor timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err
echo "bench exit $?" >> gpurun_out/bench_s2.err
DDD1D_TC_SLOTS=1 DDD1D_ENGINE=tensor timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err
DDD1D_TC_DEBUG=1 DDD1D_ENGINE=tensor timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_nomma.json 2> gpurun_out/bench_nomma.err
for w in c3 c4; do
  DDD1D_ENGINE=tensor timeout -k 10 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_s2_$w.json 2> gpurun_out/bench_s2_$w.err
done
echo done
