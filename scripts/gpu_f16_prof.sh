#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:tc_row_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_tc_f16 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_tc_f16.log 2>&1
echo done
