#!/bin/bash
# Long-horizon margins (scripts/horizon_table.py) and a C2 bench line per library variant.
# Usage: gpurun -- bash scripts/gpu_horizon_ab.sh <tag> variant...
cd "$(dirname "$0")/.."
PKG=data-driven-discretization-1d_b200
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
cp $PKG/libddd1d.so /tmp/libddd1d_keep.so
for v in "$@"; do
  cp $PKG/variants/libddd1d_$v.so $PKG/libddd1d.so
  echo "== $v"
  timeout 600 python scripts/horizon_table.py ${WLS:-c2 c3 c4} 2>&1 | tee $out/horizon_$v.txt | tail -4
  timeout -k 10 200 python bench.py --workload c2 --steps 5 --warmup 3 --rk-steps 100 --no-cpu --extra "" > $out/ab_${v}_c2.json 2> $out/ab_${v}_c2.err
  python -c "
import json; d=json.loads(open('$out/ab_${v}_c2.json').read().strip().splitlines()[-1]); print('$v', 'c2', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'])"
done
cp /tmp/libddd1d_keep.so $PKG/libddd1d.so
