#!/bin/bash
# fp16x2-plane tensor engine vs TF32 planes: tests on the default (fp16), then benches on both.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_tensor.py -q --timeout 300 -p no:cacheprovider > gpurun_out/tc_tests.log 2>&1
echo "tc tests exit $?" >> gpurun_out/tc_tests.log
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "parity exit $?" >> gpurun_out/pytest_gpu.log
DDD1D_ENGINE=tensor timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err
echo "bench f16 exit $?" >> gpurun_out/bench_f16.err
DDD1D_TC_F16=0 DDD1D_ENGINE=tensor timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
echo "bench tf32 exit $?" >> gpurun_out/bench_tf32.err
for w in c3 c4; do
  DDD1D_ENGINE=tensor timeout -k 10 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_f16_$w.json 2> gpurun_out/bench_f16_$w.err
done
timeout -k 10 300 python scripts/tc_precision.py > gpurun_out/tc_precision_f16.txt 2>&1
DDD1D_TC_F16=0 timeout -k 10 300 python scripts/tc_precision.py > gpurun_out/tc_precision_tf32.txt 2>&1
echo done
