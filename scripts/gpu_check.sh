#!/bin/bash
# One bounded pass on the GPU box: sanity -> parity tests -> smoke -> bench -> ncu.
# Everything lands in gpurun_out/.  Usage: scripts/gpu_check.sh [quick|full]
MODE=${1:-full}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt

# 1. sanity: one tiny launch per kernel flavour, hard-bounded (a hung mbarrier wait must not eat the box)
timeout -k 5 180 python - > gpurun_out/sanity.log 2>&1 <<'PY'
import numpy as np, torch, time
t0 = time.time()
import ddd1d_b200 as ddd
eq = ddd.equations.BurgersEquation(64)
s = ddd.integrate.BatchIntegrator.baseline([eq], 1)
y = s.integrate(np.zeros((1, 64), np.float32), 0.0, 1e-2, 10, 10)
torch.cuda.synchronize()
print('stencil kernel ok', float(y.abs().max()), time.time() - t0)
PY
SANITY=$?
echo "sanity exit $SANITY" >> gpurun_out/sanity.log
if [ $SANITY -ne 0 ]; then
  echo "sanity failed with bulk copy; retrying with DDD1D_NO_BULK_COPY=1" >> gpurun_out/sanity.log
  export DDD1D_NO_BULK_COPY=1
fi

# 2. parity tests through the C ABI
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log

# 3. smoke
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log

# 4. bench (N=1)
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err

if [ "$MODE" = "full" ]; then
  # 5. launch list (cold cache, serialised: shares only)
  timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --rk-steps 20 \
      > gpurun_out/ncu_launches.log 2>&1
  # 6. one full capture of the dominant kernel
  timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:row_kernel -s 3 -c 1 \
      -f -o gpurun_out/prof_row_kernel python bench.py --steps 2 --warmup 3 --no-cpu --rk-steps 10 \
      > gpurun_out/ncu_full.log 2>&1
  timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
fi
echo done
