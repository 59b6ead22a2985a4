"""bench.py -- grid-point-steps/sec of the fused integrator on BASELINE config 2
(Burgers, learned conv-net coefficients, N=256, batch 4096 per GPU).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One bench "step" = one snapshot interval of the hot path over the whole batch:
`--rk-steps` (default 50 = the reference's default output spacing 0.05 / dt 1e-3)
Bogacki-Shampine RK3 steps (3 RHS evaluations each) + one snapshot written.
metric = batch * N * rk_steps / time.  Prints ONE JSON line (rank 0).

Timing: CUDA events on the launching stream around each step, L2 flushed (256 MiB
write) between steps outside the event pair, warm-up first, max over ranks.
`value`: inputs already resident in HBM.  `e2e`: ddd1d_integrate_host with pinned
HOST buffers, host<->device copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, variant, N, per-GPU batch, dt, mode)
    'c2': ('burgers', 'plain', 256, 4096, 1e-3, 'learned'),
    'c3': ('kdv', 'plain', 128, 4096, 2.5e-5, 'learned'),
    'c4': ('ks', 'plain', 512, 4096, 1e-5, 'learned'),
    'c5': ('burgers', 'godunov', 2048, 8192, 1e-4, 'weno'),
    'c1b': ('burgers', 'plain', 64, 65536, 1e-2, 'fd'),
}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # nominal FFMA peak of a B200 (148 SMs)


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      return json.load(f), 'measured'
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


def flops_per_gps(kind, mode):
  """Algorithmic FLOPs per grid-point-step (MAC = 2, 3 RHS per step)."""
  if mode == 'learned':
    mac = {'burgers': 160 + 5120 + 1440 + 63 + 14, 'kdv': 160 + 5120 + 1280 + 56 + 14,
           'ks': 160 + 5120 + 1760 + 77 + 21}[kind]
    return 3 * 2 * mac
  return 460.0 if mode == 'weno' else 75.0


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons sampled during the timed region."""
  QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index, self.lines, self.proc = index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
           '--format=csv,noheader,nounits', '-lms', '25'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      threading.Thread(target=self._read, daemon=True).start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    sm, mx, reasons = [], None, set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for line in self.lines:
      parts = [p.strip() for p in line.split(',')]
      if len(parts) < 7:
        continue
      try:
        sm.append(float(parts[0]))
        mx = float(parts[1])
      except ValueError:
        continue
      for name, flag in zip(names, parts[3:7]):
        if flag.lower().startswith('active'):
          reasons.add(name)
    busy = sorted(sm)[len(sm) // 4:] if sm else []     # drop idle samples at the edges
    return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': mx,
            'reasons': sorted(reasons), 'samples': len(sm)}


def usable_cores():
  """Host threads this process may really use: min(cpu_count, affinity mask, cgroup quota)."""
  n = os.cpu_count() or 1
  try:
    n = min(n, len(os.sched_getaffinity(0)))
  except (AttributeError, OSError):
    pass
  for path in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
    try:
      with open(path) as f:
        fields = f.read().split()
      if path.endswith('cpu.max'):
        if fields[0] != 'max':
          n = min(n, max(1, int(int(fields[0]) / int(fields[1]))))
      else:
        quota = int(fields[0])
        if quota > 0:
          with open('/sys/fs/cgroup/cpu/cpu.cfs_period_us') as g:
            n = min(n, max(1, quota // int(g.read().split()[0])))
    except (OSError, ValueError, IndexError):
      continue
  return n


def build_case(workload, batch, seed_offset=0):
  import ddd1d_b200 as ddd
  from oracle import pde_oracle as O     # weights only: deterministic Glorot init shared with the tests
  kind, variant, n, _, dt, mode = WORKLOADS[workload]
  reg = {'plain': ddd.equations.EQUATION_TYPES, 'conservative': ddd.equations.CONSERVATIVE_EQUATION_TYPES,
         'godunov': ddd.equations.FLUX_EQUATION_TYPES}[variant]
  eqs = [reg[kind](n, random_seed=seed_offset + s) for s in range(batch)]
  if mode == 'learned':
    hp = ddd.training.create_hparams(kind, conservative=False, resample_factor=1,
                                     equation_kwargs=json.dumps({'num_points': n}))
    weights = synthetic_weights(kind, n)
    integrator = ddd.integrate.BatchIntegrator.learned(eqs, hp, weights)
  elif mode == 'weno':
    integrator = ddd.integrate.BatchIntegrator.weno(eqs)
  else:
    integrator = ddd.integrate.BatchIntegrator.baseline(eqs, 1)
  return integrator, dt, n


def synthetic_weights(kind, n):
  """Seeded Glorot-uniform kernels (tf.layers.conv1d's default initialiser), zero
  biases, last layer scaled by 1e-2 so the scheme stays a small perturbation of the
  7-point FD bias and integrates stably."""
  import math
  c_out = {'burgers': 9, 'kdv': 8, 'ks': 11}[kind]
  rs = np.random.RandomState(0)
  shapes = [(5, 1, 32), (5, 32, 32), (5, 32, c_out)]
  out = []
  for i, (k, cin, cout) in enumerate(shapes):
    limit = math.sqrt(6.0 / (k * cin + k * cout))
    w = rs.uniform(-limit, limit, size=(k, cin, cout))
    if i == len(shapes) - 1:
      w = w * 1e-2
    out.append((w.astype(np.float32), np.zeros(cout, np.float32)))
  return out


def initial_rows(batch, n, seed, workload=None):
  if workload == 'c1b':
    return np.zeros((batch, n), np.float32)       # BurgersEquation.initial_value() (equations.py:256-257)
  rs = np.random.RandomState(seed)
  x = 2 * np.pi * np.arange(n) / n
  rows = np.zeros((batch, n))
  for m in range(1, 4):
    rows += rs.randn(batch, 1) * np.sin(m * x + 2 * np.pi * rs.rand(batch, 1)) / m
  return (0.5 * rows).astype(np.float32)


# ------------------------------------------------------------------------------------
# CPU arm: the oracle restatement driven exactly like the reference's integrate.odeint
# ------------------------------------------------------------------------------------
def _cpu_one_sample(args):
  workload, seed, rk_steps = args
  from oracle import pde_oracle as O
  try:
    from threadpoolctl import threadpool_limits
    limiter = threadpool_limits(limits=1)
  except ImportError:
    limiter = None
  import scipy.integrate
  kind, variant, n, _, dt, mode = WORKLOADS[workload]
  eq = O.EquationSpec(kind, variant, num_points=n, random_seed=seed)
  if mode == 'learned':
    diff = O.ModelDifferentiator(eq, O.NetSpec(), synthetic_weights(kind, n))
  elif mode == 'weno':
    diff = O.WENODifferentiator(eq)
  else:
    diff = O.PolynomialDifferentiator(eq, 1)
  y0 = initial_rows(1, n, seed, workload)[0].astype(np.float64)
  t_end = rk_steps * dt
  # SciPy RK23 with the controller pinned at max_step (the reference's regime, integrate.py:154-155)
  sol = scipy.integrate.solve_ivp(diff, (0.0, t_end), y0, t_eval=[0.0, t_end], max_step=dt, method='RK23')
  del limiter
  return int(sol.nfev)


def cpu_reference_rate(workload, rk_steps, samples, processes, pool=None):
  """grid-point-steps/sec of the CPU path: `samples` independent samples, one
  solve_ivp each (the reference's execution model), on `processes` worker processes.
  The pool (if any) is created by the caller so its start-up is not timed."""
  kind, variant, n, _, dt, mode = WORKLOADS[workload]
  jobs = [(workload, s, rk_steps) for s in range(samples)]
  t0 = time.perf_counter()
  if pool is not None:
    nfev = pool.map(_cpu_one_sample, jobs, chunksize=1)
  else:
    nfev = [_cpu_one_sample(j) for j in jobs]
  elapsed = time.perf_counter() - t0
  return samples * n * rk_steps / elapsed, elapsed, nfev


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import multiprocessing as mp
  cores = usable_cores()
  kind, variant, n, batch, dt, mode = WORKLOADS[args.workload]
  samples = max(4 * cores, 16)
  rk = args.rk_steps
  rates = []
  with mp.get_context('spawn').Pool(cores) as pool:
    pool.map(_cpu_one_sample, [(args.workload, s, 2) for s in range(cores)])   # import + warm the workers
    for i in range(args.warmup + args.steps):
      rate, elapsed, _ = cpu_reference_rate(args.workload, rk, samples, cores, pool)
      if i >= args.warmup:
        rates.append((rate, elapsed))
  value = float(np.mean([r for r, _ in rates]))
  ms = float(np.mean([e for _, e in rates]) * 1e3)
  sample = ('%d samples x %d RK3 steps of %s %s N=%d per bench step (a bounded sample of the %d-row batch), one '
            'scipy solve_ivp(RK23, max_step=dt) per sample over a %d-process pool, BLAS 1 thread per worker '
            '(oracle port of integrate.odeint: tensorflow<2 is not installable offline so the literal TF graph '
            'cannot run)' % (samples, rk, kind, mode, n, batch, cores))
  cfg = workload_config(args, batch)
  cfg['cpu_samples_per_step'] = samples
  line = {
      'impl': 'reference', 'metric': 'grid-point-steps/sec', 'value': value, 'unit': 'grid-point-steps/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic', 'config': cfg,
      'cpu_baseline': {'value': value, 'unit': 'grid-point-steps/s', 'cores': cores, 'kind': 'port',
                       'sample': sample},
      'e2e': {'value': value, 'unit': 'grid-point-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  emit(line)


def workload_config(args, batch):
  kind, variant, n, _, dt, mode = WORKLOADS[args.workload]
  return {
      'workload': '%s: %s %s %s coefficients, N=%d, batch=%d per GPU, dt=%g, Bogacki-Shampine RK3'
                  % (args.workload, kind, variant, mode, n, batch, dt),
      'global_batch': batch * args.gpus, 'num_points': n, 'rk_steps_per_step': args.rk_steps,
      'parallelism': 'batch-sharded x%d, no data-path collective' % args.gpus,
      'l2': 'flushed between timed steps (256 MiB write)',
      'weights': 'seeded Glorot-uniform, last layer x1e-2, zero biases (random init of the reference architecture)',
      'arithmetic': 'float32 state I/O, float64 RK accumulation; conv stack FP32 FFMA, or tcgen05 kind::f16 on fp16 hi/lo planes with FP32 accumulate (FP32-faithful, tests/test_gpu_tensor.py)',
  }


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------
def run_ours(args):
  import torch
  import ddd1d_b200 as ddd
  rank, world, local = ddd.distributed.init_from_env()
  if world != args.gpus and world > 1:
    raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
  dev = torch.device('cuda', local)
  torch.cuda.set_device(dev)
  kind, variant, n, batch, dt, mode = WORKLOADS[args.workload]
  if args.batch:
    batch = args.batch
  integrator, dt, n = build_case(args.workload, batch, seed_offset=rank * batch)
  solver = integrator.solver
  rk = args.rk_steps
  u0 = torch.as_tensor(initial_rows(batch, n, seed=1000 + rank, workload=args.workload)).to(dev)
  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
  stream = torch.cuda.current_stream(dev)

  def barrier():
    if world > 1:
      torch.distributed.barrier()
    torch.cuda.synchronize(dev)

  def one_step(t0):
    return solver.integrate(u0, t0, dt, rk, rk, 'rk3')

  # ---- warm-up ----
  for i in range(max(args.warmup, 3)):
    out = one_step(0.0)
  torch.cuda.synchronize(dev)
  assert os.environ.get("DDD1D_TC_DEBUG") or torch.isfinite(out).all(), "bench workload diverged"

  # ---- timed: device-resident inputs ----
  sampler = ClockSampler(local)
  starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
  stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
  launches0 = solver.launch_count()
  barrier()
  sampler.start()
  wall0 = time.perf_counter()
  for i in range(args.steps):
    flush.fill_(float(i))                     # evict L2 (126 MB) between timed steps
    starts[i].record(stream)
    out = one_step(i * rk * dt)
    stops[i].record(stream)
  barrier()
  wall = time.perf_counter() - wall0
  clocks = sampler.stop()
  launches = solver.launch_count() - launches0
  per_step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
  local_ms = float(np.sum(per_step_ms))
  total_ms = ddd.distributed.max_over_ranks(local_ms)
  ms_per_step = total_ms / args.steps
  units_per_step = batch * world * n * rk
  value = units_per_step / (ms_per_step * 1e-3)

  # ---- timed: end to end through the C ABI with pinned host buffers ----
  host_in = torch.as_tensor(initial_rows(batch, n, seed=1000 + rank, workload=args.workload)).pin_memory()
  host_out = torch.empty((1, batch, n), dtype=torch.float32).pin_memory()
  host_bad = torch.empty(batch, dtype=torch.int32).pin_memory()
  import ctypes
  lib, handle = solver._lib, solver._handle

  def e2e_step(t0):
    ddd._lib.check(lib.ddd1d_integrate_host(handle, float(t0), float(dt), rk, rk, 0, host_in.data_ptr(),
                                            host_out.data_ptr(), host_bad.data_ptr(), batch, 0), handle)
    return float(host_out[0, 0, 0])           # read the result on the host

  for i in range(2):
    e2e_step(0.0)
  barrier()
  e0 = time.perf_counter()
  for i in range(args.steps):
    e2e_step(i * rk * dt)
  torch.cuda.synchronize(dev)
  e2e_local = time.perf_counter() - e0
  e2e_total = ddd.distributed.max_over_ranks(e2e_local)
  e2e_value = units_per_step * args.steps / e2e_total

  # the one collective of the path: gather the final snapshots of all shards (NCCL over NVLink), untimed
  gather_ms = None
  if world > 1:
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    g0.record(stream)
    full = ddd.distributed.gather_snapshots(out, batch * world, sample_axis=1)
    g1.record(stream)
    torch.cuda.synchronize(dev)
    assert tuple(full.shape) == (1, batch * world, n)
    gather_ms = ddd.distributed.max_over_ranks(g0.elapsed_time(g1))

  if rank != 0:
    if world > 1:
      torch.distributed.barrier()
      torch.distributed.destroy_process_group()
    return

  peaks, peak_kind = measured_peaks()
  kernel_ms = float(np.mean(per_step_ms))       # one kernel launch per step on this stream
  gps_kernel = batch * n * rk / (kernel_ms * 1e-3)
  achieved_gbs = 8.0 * gps_kernel / 1e9          # 8 algorithmic bytes per grid-point-step
  fl = flops_per_gps(kind, mode)
  shape = solver.launch_shape(batch)
  engine = solver.engine() if mode == 'learned' else 'ffma'
  kernel_name = 'ddd1d::tc::tc_row_kernel' if engine == 'tensor' else 'ddd1d::row_kernel<%s>' % mode
  hbm = {'bound': 'hbm', 'achieved': achieved_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
         'frac': achieved_gbs / peaks['hbm_gbs'], 'traffic': None, 'peak_kind': peak_kind,
         'algorithmic_bytes_per_launch': 8.0 * batch * n * rk,
         'note': 'rows stay on chip for all RK steps of a launch (ncu: DRAM traffic ~7% of the algorithmic '
                 'bytes), so HBM is idle by design; the binding resource is on-chip'}
  traffic = None
  tpath = os.path.join(ROOT, 'profiles', 'r01', 'traffic.json')
  if os.path.exists(tpath) and args.workload == 'c2' and rk == 50 and batch == 4096:
    with open(tpath) as f:
      traffic = json.load(f).get(kernel_name, {}).get('dram_bytes_per_launch')   # ncu --set full, same launch shape
  hbm['traffic'] = traffic
  achieved_tf = fl * gps_kernel / 1e12
  if engine == 'tensor':
    # the conv stack runs as kind::f16 MMAs (fp16 hi/lo planes, FP32 accumulate): the dense f16/bf16 rate is the peak
    f16_peak = peaks['bf16_tflops_sustained' if 'bf16_tflops_sustained' in peaks else 'bf16_tflops']
    roofline = {'bound': 'tensor', 'achieved': achieved_tf, 'peak': f16_peak, 'unit': 'TFLOP/s',
                'frac': achieved_tf / f16_peak, 'traffic': traffic, 'peak_kind': peak_kind,
                'note': 'achieved = algorithmic FP32-equivalent FLOPs (%d per grid-point-step); peak = measured sustained '
                        'dense bf16/f16 cuBLAS rate.  The FP32-faithful fp16 hi/lo split executes 3x the conv FLOPs on '
                        'the tensor pipe (hi*Wh, hi*Wl, lo*Wh) at N padded to 32/16 and every MMA streams its operands '
                        'from shared memory (88 clk per K16 step for 48 clk of math), so the MMA stream alone needs 7.7 ms '
                        'of the launch; the CUDA-core side (stage values, first layer, splits, epilogues) alone needs 9.8 ms '
                        'and the two overlap by about half (profiles/r01/README.md)' % fl}
  else:
    roofline = dict(hbm)
  roofline.update({'kernel': kernel_name, 'kernel_ms': kernel_ms})
  line = {
      'metric': 'grid-point-steps/sec', 'value': value, 'unit': 'grid-point-steps/s',
      'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic', 'config': workload_config(args, batch),
      'clocks': clocks,
      'e2e': {'value': e2e_value, 'unit': 'grid-point-steps/s',
              'h2d_bytes_per_step': batch * n * 4, 'd2h_bytes_per_step': batch * n * 4 + batch * 4},
      'gpu_launches': int(launches),
      'engine': engine,
      'roofline': roofline,
      'roofline_hbm': hbm,
      'compute': {'bound': 'fp32-ffma', 'achieved_tflops': achieved_tf,
                  'peak_tflops': FP32_PEAK_TFLOPS, 'frac': achieved_tf / FP32_PEAK_TFLOPS,
                  'flops_per_grid_point_step': fl, 'peak_kind': 'nominal 148 SM x 128 FMA x 1.965 GHz',
                  'note': 'FP32-equivalent algorithmic FLOPs against the CUDA-core peak (can exceed 1 on the tensor engine)'},
      'launch': shape, 'wall_s': wall, 'final_gather_ms': gather_ms,
  }
  if not args.no_cpu:
    samples = 160                             # ~10-15 s of single-core work (the bounded sample of the batch)
    rk_cpu = rk
    _cpu_one_sample((args.workload, 0, 2))     # imports, first-call costs
    rate, elapsed, nfev = cpu_reference_rate(args.workload, rk_cpu, samples, 1)
    line['cpu_baseline'] = {
        'value': rate, 'unit': 'grid-point-steps/s', 'cores': 1, 'kind': 'port',
        'sample': '%d samples x %d RK3 steps, N=%d, scipy solve_ivp(RK23, max_step=dt) on the oracle port, '
                  '1 process, BLAS limited to 1 thread, %.1f s' % (samples, rk_cpu, n, elapsed)}
  emit(line)
  if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


_RESULT_FD = None


def emit(line):
  """Print the one JSON line on the real stdout (see main: fd 1 is pointed at stderr while
  libraries such as NCCL chat)."""
  data = (json.dumps(line) + '\n').encode()
  if _RESULT_FD is None:
    sys.stdout.write(data.decode())
    sys.stdout.flush()
  else:
    os.write(_RESULT_FD, data)


def main():
  global _RESULT_FD
  # keep stdout for the single JSON line: NCCL prints its version banner on fd 1
  sys.stdout.flush()
  _RESULT_FD = os.dup(1)
  os.dup2(2, 1)
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
  ap.add_argument('--rk-steps', type=int, default=50, dest='rk_steps')
  ap.add_argument('--batch', type=int, default=0, help='per-GPU batch override')
  ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
