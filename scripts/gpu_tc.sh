#!/bin/bash
# Tensor-engine bring-up pass: probe first (bounded), then the tensor tests, then benches on both engines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_tensor.py -q --timeout 120 -p no:cacheprovider -k probe > gpurun_out/tc_probe.log 2>&1
echo "probe exit $?" >> gpurun_out/tc_probe.log
timeout -k 10 900 python -m pytest tests/test_gpu_tensor.py -q --timeout 300 -p no:cacheprovider -k "not probe" > gpurun_out/tc_tests.log 2>&1
echo "tc tests exit $?" >> gpurun_out/tc_tests.log
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "parity exit $?" >> gpurun_out/pytest_gpu.log
DDD1D_ENGINE=tensor timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err
echo "bench tensor exit $?" >> gpurun_out/bench_tensor.err
DDD1D_ENGINE=ffma timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_ffma.json 2> gpurun_out/bench_ffma.err
echo "bench ffma exit $?" >> gpurun_out/bench_ffma.err
cat /sys/fs/cgroup/cpu.max > gpurun_out/cpu_max.txt 2>&1
python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count())" >> gpurun_out/cpu_max.txt 2>&1
echo done
