#!/bin/bash
# A/B of library variants incl. the MMA-stream-only timing mode
cd "$(dirname "$0")/.."
PKG=data-driven-discretization-1d_b200
mkdir -p gpurun_out
cp $PKG/libddd1d.so /tmp/libddd1d_keep.so
for v in "$@"; do
  cp $PKG/variants/libddd1d_$v.so $PKG/libddd1d.so
  timeout -k 10 300 python -m pytest tests/test_gpu_tensor.py -q -x --timeout 200 -p no:cacheprovider 2>&1 | tail -1
  for mode in 0 64; do
  for w in c2 c3 c4; do
    DDD1D_TC_DEBUG=$mode timeout -k 10 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/ab_${v}_$w.json 2> gpurun_out/ab_${v}_$w.err
    python -c "
import json; d=json.loads(open('gpurun_out/ab_${v}_$w.json').read().strip().splitlines()[-1]); print('$v', 'debug=$mode', '$w', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'])"
  done
  done
done
cp /tmp/libddd1d_keep.so $PKG/libddd1d.so
