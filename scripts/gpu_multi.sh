#!/bin/bash
# 2-GPU pass: torchrun bench (ours + reference arm), then the non-headline workloads on GPU 0.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 exit $?" >> gpurun_out/bench_n2.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "ref n2 exit $?" >> gpurun_out/bench_ref_n2.err
for w in c5 c1b; do
  timeout -k 10 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "$w exit $?" >> gpurun_out/bench_$w.err
done
DDD1D_ENGINE=ffma timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_ffma.json 2> gpurun_out/bench_ffma.err
echo done
