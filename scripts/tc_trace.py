"""Debug: run one C2-shaped launch with the event trace on (DDD1D_TC_TRACE) and keep the raw trace."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/tc_trace.bin'
os.environ['DDD1D_TC_TRACE'] = out
import torch
import bench
integrator, dt, n = bench.build_case('c2', 4096)
u0 = torch.as_tensor(bench.initial_rows(4096, n, seed=1000, workload='c2')).cuda()
for _ in range(2):
  y = integrator.solver.integrate(u0, 0.0, dt, 50, 50, 'rk3')
torch.cuda.synchronize()
print('trace written', out, os.path.getsize(out))
