#!/bin/bash
# A/B of library variants in data-driven-discretization-1d_b200/variants/ on one box: tensor-engine parity tests
# per variant, then bench lines of WLS (default c2 c3 c4 c2s), two rounds, interleaved.
# Usage: gpurun -- bash scripts/gpu_ab2.sh <tag> variant...
cd "$(dirname "$0")/.."
PKG=data-driven-discretization-1d_b200
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
cp $PKG/libddd1d.so /tmp/libddd1d_keep.so
for rep in 1 2; do
for v in "$@"; do
  cp $PKG/variants/libddd1d_$v.so $PKG/libddd1d.so
  if [ $rep = 1 ]; then
    timeout -k 10 400 python -m pytest ${TESTS:-tests/test_gpu_tensor.py} -q -x --timeout 300 -p no:cacheprovider > $out/pytest_$v.log 2>&1
    echo "$v pytest: $(tail -1 $out/pytest_$v.log)"
  fi
  for w in ${WLS:-c2 c3 c4 c2s}; do
    timeout -k 10 200 python bench.py --workload $w --steps 5 --warmup 3 --rk-steps ${RK:-100} --no-cpu --extra "" > $out/ab_${v}_${w}_$rep.json 2> $out/ab_${v}_${w}_$rep.err
    python -c "
import json; d=json.loads(open('$out/ab_${v}_${w}_$rep.json').read().strip().splitlines()[-1]); print('$v', '$w', '%.3f ms'%d['ms_per_step'], '%.3e'%d['value'], d.get('parity_check',{}).get('rel_err'))"
  done
done
done
cp /tmp/libddd1d_keep.so $PKG/libddd1d.so
