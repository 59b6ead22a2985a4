#!/bin/bash
# Round evidence on one B200: parity tests, smoke, the driver's bench command, the reference arm, variant engines,
# the ncu launch list of the same command and one `ncu --set full` capture per top kernel.
# Usage: gpurun -- bash scripts/gpu_evidence.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-evidence}
out=gpurun_out/$tag
mkdir -p $out
timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > $out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/pytest_gpu.log; tail -2 $out/pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log; tail -1 $out/smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
python bench.py --engine tensor_f16x2 --steps 20 --warmup 5 --no-cpu --extra '' > $out/bench_n1_f16x2.json 2> $out/bench_n1_f16x2.err
python bench.py --engine ffma --steps 4 --warmup 3 --rk-steps 100 --no-cpu --extra '' > $out/bench_n1_ffma.json 2> $out/bench_n1_ffma.err
for w in c3 c4; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --extra '' > $out/bench_$w.json 2> $out/bench_$w.err
done
python bench.py --workload c2s --steps 5 --warmup 3 --rk-steps 50 --no-cpu --extra '' > $out/bench_c2s.json 2> $out/bench_c2s.err
python bench.py --workload c2s --engine ffma --steps 4 --warmup 3 --rk-steps 50 --no-cpu --extra '' > $out/bench_c2s_ffma.json 2> $out/bench_c2s_ffma.err
# every launch of the bench command with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --extra '' > $out/launches_c2.log 2>&1
prof() {   # name, kernel regex, bench args...: capture, summarise on the box (gpurun_out is capped at 64 MiB)
  name=$1; regex=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$regex -s 3 -c 1 -o /tmp/prof_$name \
      python bench.py --steps 2 --warmup 3 --no-cpu --extra '' "$@" > $out/prof_$name.log 2>&1
  python scripts/ncu_summary.py /tmp/prof_$name.ncu-rep $out/${name}_ncu_full.json >> $out/prof_$name.log 2>&1
  if [ "$name" = tc_c2 ]; then cp /tmp/prof_$name.ncu-rep $out/; fi
}
prof tc_c2 tc_row_kernel --workload c2 --rk-steps 20
prof tc_c3 tc_row_kernel --workload c3 --rk-steps 20
prof tc_c2s tc_row_kernel --workload c2s --rk-steps 20
prof weno_c5 weno_block_kernel --workload c5 --rk-steps 5
prof warp_c1b warp_row_kernel --workload c1b --rk-steps 20
for dbg in 1 64; do
  DDD1D_TC_DEBUG=$dbg python bench.py --steps 5 --warmup 3 --rk-steps 50 --no-cpu --extra '' > $out/c2_d$dbg.json 2> $out/c2_d$dbg.err
done
ls -la $out | head -60
