#!/bin/bash
# Evidence pass for the current tensor engine: parity tests -> smoke -> bench -> launch list -> ncu full -> reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_c2_tensor.csv python bench.py --steps 3 --warmup 3 --no-cpu \
    > gpurun_out/ncu_launches.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:tc_row_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_tc_f16 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
for w in c3 c4; do
  timeout -k 10 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo done
