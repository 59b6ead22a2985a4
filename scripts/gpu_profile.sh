#!/bin/bash
# Evidence pass: launch list + full ncu capture of the dominant kernel at the bench's launch shape.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_tensor.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launches_tensor.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:tc_row_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_tc_final python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_tc_final.log 2>&1
DDD1D_ENGINE=ffma timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:row_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_ffma_final python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_ffma_final.log 2>&1
echo done
