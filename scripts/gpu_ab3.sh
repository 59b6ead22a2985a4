#!/bin/bash
cd "$(dirname "$0")/.."
PKG=data-driven-discretization-1d_b200
mkdir -p gpurun_out
cp $PKG/libddd1d.so /tmp/libddd1d_keep.so
W=$1; shift
for rep in 1 2; do
for v in "$@"; do
  cp $PKG/variants/libddd1d_$v.so $PKG/libddd1d.so
  timeout -k 10 200 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$W', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'], d['launch'])"
done
done
cp /tmp/libddd1d_keep.so $PKG/libddd1d.so
