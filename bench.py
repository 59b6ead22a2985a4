"""bench.py -- grid-point-steps/sec of the fused integrator on BASELINE config 2
(Burgers, learned conv-net coefficients, N=256, batch 4096 per GPU, 10 000 RK3 steps).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One bench "step" = `--rk-steps` (default 500) Bogacki-Shampine RK3 steps (3 right-hand sides each) of
the whole batch in ONE kernel launch + one snapshot written.  The timed steps form ONE trajectory: step
i starts from step i-1's snapshot at t = i * rk_steps * dt, so the driver's `--steps 20` IS the
configuration's 10 000 steps (warm-up launches start from the same u0 and are discarded).
metric = batch * N * rk_steps / time.  Prints ONE JSON line (rank 0).

Timing: CUDA events on the launching stream around each step, L2 flushed (256 MiB write) between
steps outside the event pair, warm-up first, max over ranks.  `value`: inputs already resident in HBM.
`e2e`: the same trajectory through the C ABI with pinned HOST buffers (ddd1d_integrate_host at N=1;
at N>1 pinned-host -> device copy, ddd1d_integrate, the NCCL all-gather of the step's snapshot -- the one
collective of the path -- and the device -> host read), every copy inside the timed region.
`parity_check`: the trajectory the timed steps produced, against the oracle (committed long-horizon
fixture when the run is the configured one, else two rows integrated live by the CPU oracle).
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

FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # nominal FFMA peak of a B200 (148 SMs)
STRONG_BATCH = 4096                                   # --scaling strong: the metric's batch over all GPUs


def workloads():
  import ddd1d_b200.workloads as wl
  return wl


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      return json.load(f), 'measured'
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


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
    sm, mx, power, reasons = [], None, [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for line in self.lines:
      parts = [p.strip() for p in line.split(',')]
      if len(parts) < 7:
        continue
      try:
        sm.append(float(parts[0]))
        mx = float(parts[1])
        power.append(float(parts[2]))
      except ValueError:
        continue
      for name, flag in zip(names, parts[3:7]):
        if flag.lower().startswith('active'):
          reasons.add(name)
    busy = sorted(sm)[len(sm) // 4:] if sm else []     # drop idle samples at the edges
    return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': mx,
            'reasons': sorted(reasons), 'samples': len(sm), 'power_w_max': max(power) if power else None}


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


def build_case(workload, batch, seed_offset=0, engine='auto'):
  import ddd1d_b200 as ddd
  wl = workloads()
  kind, variant, n, _, dt, mode = wl.WORKLOADS[workload]
  reg = {'plain': ddd.equations.EQUATION_TYPES, 'conservative': ddd.equations.CONSERVATIVE_EQUATION_TYPES,
         'godunov': ddd.equations.FLUX_EQUATION_TYPES}[variant]
  eqs = [reg[kind](n, random_seed=seed_offset + s) for s in range(batch)]
  if mode == 'learned':
    hp = ddd.training.create_hparams(kind, conservative=False, resample_factor=1,
                                     equation_kwargs=json.dumps({'num_points': n}))
    integrator = ddd.integrate.BatchIntegrator.learned(eqs, hp, wl.synthetic_weights(kind), engine=engine)
  elif mode == 'weno':
    integrator = ddd.integrate.BatchIntegrator.weno(eqs)
  else:
    integrator = ddd.integrate.BatchIntegrator.baseline(eqs, 1)
  return integrator, dt, n


def bench_rows(workload, batch, rank):
  """Initial rows of one rank.  Rank 0 of the configured batch integrates exactly the rows of the
  long-horizon parity fixture (tests/golden/long_horizon.npz)."""
  wl = workloads()
  kind, _, n, default_batch, _, mode = wl.WORKLOADS[workload]
  if rank == 0 and batch == default_batch and workload in ('c2', 'c3', 'c4'):
    return wl.horizon_rows(workload)
  return wl.initial_rows(batch, n, wl.HORIZON_SEED + rank, workload)


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
  wl = workloads()
  kind, variant, n, _, dt, mode = wl.WORKLOADS[workload]
  eq = O.EquationSpec(kind, variant, num_points=n, random_seed=seed)
  if mode == 'learned':
    diff = O.ModelDifferentiator(eq, O.NetSpec(), wl.synthetic_weights(kind))
  elif mode == 'weno':
    diff = O.WENODifferentiator(eq)
  else:
    diff = O.PolynomialDifferentiator(eq, 1)
  y0 = wl.initial_rows(1, n, seed, workload)[0].astype(np.float64)
  t_end = rk_steps * dt
  # SciPy RK23 with the controller pinned at max_step (the reference's regime, integrate.py:154-155)
  sol = scipy.integrate.solve_ivp(diff, (0.0, t_end), y0, t_eval=[0.0, t_end], max_step=dt, method='RK23')
  del limiter
  return int(sol.nfev)


def cpu_reference_rate(workload, rk_steps, samples, pool=None):
  """grid-point-steps/sec of the CPU path: `samples` independent samples, one solve_ivp each (the
  reference's execution model).  The pool (if any) is created by the caller so its start-up is not timed."""
  n = workloads().WORKLOADS[workload][2]
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
  kind, variant, n, batch, dt, mode = workloads().WORKLOADS[args.workload]
  batch = per_gpu_batch(args, batch)
  samples = max(cores, 8)
  rk = args.rk_steps
  rates = []
  with mp.get_context('spawn').Pool(cores) as pool:
    pool.map(_cpu_one_sample, [(args.workload, s, 2) for s in range(cores)])   # import + warm the workers
    for i in range(args.warmup + args.steps):
      rate, elapsed, _ = cpu_reference_rate(args.workload, rk, samples, pool)
      if i >= args.warmup:
        rates.append((rate, elapsed))
  value = float(np.mean([r for r, _ in rates]))
  ms = float(np.mean([e for _, e in rates]) * 1e3)
  sample = ('%d samples x %d RK3 steps of %s %s N=%d per bench step (a bounded sample of the %d-row batch; rows '
            'are independent, so the rate per sample is the rate of the batch), one scipy solve_ivp(RK23, '
            'max_step=dt) per sample over a %d-process pool, BLAS 1 thread per worker (oracle port of '
            'integrate.odeint: tensorflow<2 is not installable offline so the literal TF graph cannot run)'
            % (samples, rk, kind, mode, n, batch, cores))
  line = {
      'impl': 'reference', 'metric': 'grid-point-steps/sec', 'value': value, 'unit': 'grid-point-steps/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
      'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic', 'config': workload_config(args, batch),
      'cpu_baseline': {'value': value, 'unit': 'grid-point-steps/s', 'cores': cores, 'kind': 'port',
                       'sample': sample},
      'e2e': {'value': value, 'unit': 'grid-point-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  emit(line)


def per_gpu_batch(args, default_batch):
  if args.batch:
    return args.batch
  if args.scaling == 'strong':
    return max(1, STRONG_BATCH // max(args.gpus, 1))
  return default_batch


def workload_config(args, batch):
  kind, variant, n, _, dt, mode = workloads().WORKLOADS[args.workload]
  return {
      'workload': '%s: %s %s %s coefficients, N=%d, batch=%d per GPU, dt=%g, Bogacki-Shampine RK3, %d steps'
                  % (args.workload, kind, variant, mode, n, batch, dt, args.steps * args.rk_steps),
      'global_batch': batch * args.gpus, 'num_points': n, 'rk_steps_per_step': args.rk_steps,
      'trajectory_steps': args.steps * args.rk_steps,
      'parallelism': 'batch-sharded x%d, no data-path collective' % args.gpus,
      'l2': 'flushed between timed steps (256 MiB write)',
      'weights': 'seeded Glorot-uniform, last layer x1e-2, zero biases (random init of the reference architecture)',
      'arithmetic': 'float32 state I/O, float64 RK accumulation; conv stack FP32 FFMA, or tcgen05 kind::f16 on fp16 hi/lo planes with FP32 accumulate (FP32-faithful, tests/test_gpu_tensor.py)',
  }


# ------------------------------------------------------------------------------------
# parity of the timed trajectory
# ------------------------------------------------------------------------------------
def parity_check(args, batch, final_rows, total_steps):
  """The state the timed steps ended in against the oracle.  `final_rows`: {row index: float32 [N]}."""
  wl = workloads()
  kind, variant, n, default_batch, dt, mode = wl.WORKLOADS[args.workload]
  rows = sorted(final_rows)
  got = np.stack([final_rows[r] for r in rows]).astype(np.float64)
  fixture = os.path.join(ROOT, 'tests', 'golden', 'long_horizon.npz')
  if (args.workload in ('c2', 'c3', 'c4') and batch == default_batch and os.path.exists(fixture)
      and total_steps % 1000 == 0 and 0 < total_steps <= wl.FULL_STEPS):
    with np.load(fixture) as f:
      picks = list(f['%s/rows' % args.workload])
      idx = [picks.index(r) for r in rows]
      want32 = f['%s/f32' % args.workload][total_steps // 1000 - 1][idx]
      want64 = f['%s/f64' % args.workload][total_steps // 1000 - 1][idx]
    scale = np.abs(want64).max()
    return {'rows': rows, 'steps': total_steps, 'against': 'tests/golden/long_horizon.npz (oracle, float64 state)',
            'rel_err': float(np.abs(got - want32).max() / scale),
            'rel_err_vs_float64_oracle': float(np.abs(got - want64).max() / scale),
            'float32_oracle_vs_float64_oracle': float(np.abs(want32 - want64).max() / scale)}
  if mode != 'learned' or total_steps > 3000:
    return {'skipped': 'no fixture for this run (%s, %d steps) and a live oracle run would take too long'
                       % (args.workload, total_steps)}
  from oracle import pde_oracle as O
  rows = rows[:2]
  eqs = [O.EquationSpec(kind, variant, num_points=n, random_seed=int(r)) for r in rows]
  rhs = O.batched_rhs(eqs, O.NetSpec(), wl.synthetic_weights(kind), mode='learned')
  u0 = bench_rows(args.workload, batch, 0)[rows]
  t0 = time.perf_counter()
  want = O.fixed_step_integrate(rhs, u0, 0.0, dt, total_steps, total_steps)[-1]
  return {'rows': rows, 'steps': total_steps, 'against': 'oracle.fixed_step_integrate, live (%.0f s)' % (time.perf_counter() - t0),
          'rel_err': float(np.abs(got[:2] - want).max() / np.abs(want).max())}


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------
def time_trajectory(solver, u0, dt, rk, steps, warmup, flush, stream, barrier, carry=True):
  """`warmup` discarded launches from u0, then `steps` launches forming one trajectory from u0 at t = 0
  (carry=False: every launch restarts from u0).  Returns (per-step ms list, final snapshot [1, batch, N], wall seconds)."""
  import torch
  for _ in range(warmup):
    out = solver.integrate(u0, 0.0, dt, rk, rk, 'rk3')
  torch.cuda.synchronize()
  starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
  stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
  state = u0
  barrier()
  wall0 = time.perf_counter()
  for i in range(steps):
    flush.fill_(float(i))                     # evict L2 (126 MB) between timed steps
    starts[i].record(stream)
    out = solver.integrate(state, i * rk * dt if carry else 0.0, dt, rk, rk, 'rk3')
    stops[i].record(stream)
    state = out[0] if carry else u0
  barrier()
  wall = time.perf_counter() - wall0
  return [s.elapsed_time(e) for s, e in zip(starts, stops)], out, wall


def quick_rate(workload, steps=4, rk=None, engine='auto', batch=None):
  """A short device-timed measurement of another workload (reported under `other_workloads`)."""
  import torch
  wl = workloads()
  kind, variant, n, default_batch, dt, mode = wl.WORKLOADS[workload]
  batch = batch or default_batch
  rk = rk or {'c5': 10}.get(workload, 50)        # c1b: 4 x 50 = config 1's 200 steps (T = 2)
  integrator, dt, n = build_case(workload, batch, engine=engine)
  solver = integrator.solver
  dev = torch.device('cuda', torch.cuda.current_device())
  u0 = torch.as_tensor(wl.initial_rows(batch, n, wl.HORIZON_SEED, workload)).to(dev)
  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
  # c1b: config 1's horizon is 200 steps for ONE seed; over 65536 seeds some first-order rows leave the finite range
  # later, so its launches restart from u0
  ms, out, _ = time_trajectory(solver, u0, dt, rk, steps, 3, flush, torch.cuda.current_stream(dev),
                               lambda: torch.cuda.synchronize(dev), carry=workload != 'c1b')
  assert torch.isfinite(out).all(), '%s diverged' % workload
  kernel_ms = float(np.mean(ms))
  gps = batch * n * rk / (kernel_ms * 1e-3)
  peaks, _ = measured_peaks()
  eng = solver.engine() if mode == 'learned' else None
  res = {'value': gps, 'unit': 'grid-point-steps/s', 'batch': batch, 'num_points': n, 'rk_steps_per_launch': rk,
         'kernel_ms': kernel_ms, 'engine': eng, 'hbm_frac': 8.0 * gps / 1e9 / peaks['hbm_gbs']}
  if eng and eng.startswith('tensor'):
    res['tensor_frac_burst'] = flops_per_gps(kind, mode) * gps / 1e12 / peaks['bf16_tflops']
  else:
    res['fp32_frac_nominal'] = flops_per_gps(kind, mode) * gps / 1e12 / FP32_PEAK_TFLOPS
  solver.close()
  return res


def run_ours(args):
  import torch
  import ddd1d_b200 as ddd
  wl = workloads()
  rank, world, local = ddd.distributed.init_from_env()
  if world != args.gpus and world > 1:
    raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
  dev = torch.device('cuda', local)
  torch.cuda.set_device(dev)
  kind, variant, n, default_batch, dt, mode = wl.WORKLOADS[args.workload]
  batch = per_gpu_batch(args, default_batch)
  integrator, dt, n = build_case(args.workload, batch, seed_offset=rank * batch, engine=args.engine)
  solver = integrator.solver
  rk = args.rk_steps
  rows0 = bench_rows(args.workload, batch, rank)
  u0 = torch.as_tensor(rows0).to(dev)
  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
  stream = torch.cuda.current_stream(dev)
  warmup = max(args.warmup, 3)

  def barrier():
    if world > 1:
      torch.distributed.barrier()
    torch.cuda.synchronize(dev)

  if world > 1:                               # connect the communicator before anything is timed
    warm = torch.zeros((1, batch, n), dtype=torch.float32, device=dev)
    for _ in range(2):
      ddd.distributed.gather_snapshots(warm, batch * world, sample_axis=1)
    torch.cuda.synchronize(dev)

  # ---- timed: device-resident inputs ----
  sampler = ClockSampler(local)
  launches0 = solver.launch_count()
  sampler.start()
  per_step_ms, out, wall = time_trajectory(solver, u0, dt, rk, args.steps, warmup, flush, stream, barrier)
  clocks = sampler.stop()
  launches = solver.launch_count() - launches0 - warmup
  assert os.environ.get('DDD1D_TC_DEBUG') or torch.isfinite(out).all(), 'bench workload diverged'
  local_ms = float(np.sum(per_step_ms))
  total_ms = ddd.distributed.max_over_ranks(local_ms)
  ms_per_step = total_ms / args.steps
  units_per_step = batch * world * n * rk
  value = units_per_step / (ms_per_step * 1e-3)
  picks = [r for r in wl.HORIZON_PICKS if r < batch]
  final_rows = {r: out[0, r].cpu().numpy() for r in picks} if rank == 0 else {}

  # ---- timed: end to end through the C ABI with pinned host buffers, the same trajectory again ----
  host_state = torch.as_tensor(rows0).pin_memory()
  host_out = torch.empty((1, batch, n), dtype=torch.float32).pin_memory()
  host_full = torch.empty((1, batch * world, n), dtype=torch.float32).pin_memory() if world > 1 else None
  host_bad = torch.empty(batch, dtype=torch.int32).pin_memory()
  lib, handle = solver._lib, solver._handle
  gather_ms = []

  def e2e_step(i, src):
    t0 = i * rk * dt
    if world == 1:
      ddd._lib.check(lib.ddd1d_integrate_host(handle, float(t0), float(dt), rk, rk, 0, src.data_ptr(),
                                              host_out.data_ptr(), host_bad.data_ptr(), batch, 0), handle)
      return host_out
    d_in = src.to(dev, non_blocking=True)                                       # H2D from pinned memory
    snap = solver.integrate(d_in, t0, dt, rk, rk, 'rk3')
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    full = ddd.distributed.gather_snapshots(snap, batch * world, sample_axis=1)   # NCCL over NVLink
    g1.record(stream)
    host_full.copy_(full, non_blocking=True)                                    # D2H: every rank reads the gathered result
    torch.cuda.synchronize(dev)
    gather_ms.append(g0.elapsed_time(g1))
    host_out.copy_(host_full[:, rank * batch:(rank + 1) * batch])               # this rank's block feeds its next step
    return host_out

  for i in range(2):
    e2e_step(0, host_state)
  del gather_ms[:]
  barrier()
  e0 = time.perf_counter()
  src = host_state
  for i in range(args.steps):
    res = e2e_step(i, src)
    _ = float(res[0, 0, 0])                     # the result is read on the host
    host_state.copy_(res[0])
    src = host_state
  torch.cuda.synchronize(dev)
  e2e_local = time.perf_counter() - e0
  e2e_total = ddd.distributed.max_over_ranks(e2e_local)
  e2e_value = units_per_step * args.steps / e2e_total
  gather = None
  if world > 1:
    g_ms = ddd.distributed.max_over_ranks(float(np.mean(gather_ms)))
    g_bytes = batch * world * n * 4
    gather = {'ms': g_ms, 'bytes_gathered_per_rank': g_bytes,
              'algbw_gbs': g_bytes / (g_ms * 1e-3) / 1e9,
              'busbw_gbs': g_bytes * (world - 1) / world / (g_ms * 1e-3) / 1e9,
              'note': 'all_gather_into_tensor of one snapshot, communicator warmed, timed on the device inside the e2e step; '
                      'a %0.1f MiB message is latency bound, not NVLink bound (peer copy peak 770 GB/s)' % (g_bytes / 2**20)}
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()   # the other ranks leave here; rank 0 does its CPU legs alone
  if rank != 0:
    return

  peaks, peak_kind = measured_peaks()
  kernel_ms = float(np.mean(per_step_ms))       # one kernel launch per step on this stream
  gps_kernel = batch * n * rk / (kernel_ms * 1e-3)
  achieved_gbs = 8.0 * gps_kernel / 1e9          # 8 algorithmic bytes per grid-point-step
  fl = flops_per_gps(kind, mode)
  shape = solver.launch_shape(batch)
  engine = solver.engine() if mode == 'learned' else 'ffma'
  kernel_name = 'ddd1d::tc::tc_row_kernel' if engine.startswith('tensor') else 'ddd1d::row_kernel<%s>' % mode
  traffic = None
  tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
  if os.path.exists(tpath):
    with open(tpath) as f:
      entry = json.load(f).get('%s/%s' % (args.workload, engine))
    if entry and entry.get('batch') == batch:   # ncu --set full, same launch shape; bytes scale with the steps of a launch
      traffic = entry['dram_bytes_per_rk_step'] * rk
  hbm = {'bound': 'hbm', 'achieved': achieved_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
         'frac': achieved_gbs / peaks['hbm_gbs'], 'traffic': traffic, 'peak_kind': peak_kind,
         'algorithmic_bytes_per_launch': 8.0 * batch * n * rk,
         'note': 'rows stay on chip for all RK steps of a launch, so HBM is idle by design; the binding resource is on-chip'}
  achieved_tf = fl * gps_kernel / 1e12
  if engine.startswith('tensor'):
    burst = peaks['bf16_tflops']
    sustained = peaks.get('bf16_tflops_sustained', burst)
    roofline = {'bound': 'tensor', 'achieved': achieved_tf, 'peak': burst, 'unit': 'TFLOP/s',
                'frac': achieved_tf / burst, 'frac_burst': achieved_tf / burst,
                'frac_sustained': achieved_tf / sustained, 'peak_sustained': sustained,
                'traffic': traffic, 'peak_kind': peak_kind,
                'note': 'achieved = algorithmic FP32-equivalent FLOPs (%d per grid-point-step) / CUDA-event launch time; '
                        'peak = measured dense bf16/f16 cuBLAS rate (burst; the sustained figure beside it).  The '
                        'FP32-faithful fp16 hi/lo split executes 3x the conv FLOPs on the tensor pipe (hi*Wh, hi*Wl, lo*Wh) '
                        'and streams every operand from shared memory' % fl}
    # What binds this kernel in practice: the tensor core's operand path.  A tcgen05.mma of M = 128, K = 16 fetches its
    # 128 x 16 fp16 A tile in 32 clocks of shared-memory wavefronts whatever N is (+ N/4 for B), so the small-N
    # MMAs of a 32-channel conv cost 32 + N/4 clocks for N/2 clocks of math.  Floor = the measured clocks of one
    # tile's MMAs per right-hand side, streamed back to back from ONE issuer (profiles/r01/tc_mma_rate.txt; 148
    # CTAs: same), x the tiles an SM processes per launch / the SM clock sampled during the run.
    step_clk = {3: {64: 88.2, 32: 79.2}, 2: {64: 48.2, 32: 40.2}, 1: {64: 40.2, 32: 39.2}}      # [prec][2 * NB]: (tap, ci-block) step
    prec = {'tensor': 3, 'tensor_f16x2': 2, 'tensor_f16': 1}.get(engine, 3)
    nl = 32 if kind == 'ks' else 16
    clk_tile_rhs = 10.0 * (step_clk[prec][64] + step_clk[prec][2 * nl])      # one hidden tensor layer + the last layer
    tiles = batch * max(n, 128) / 128.0 if n >= 128 else batch * n / 128.0   # 128-position tiles in the batch
    sm_mhz = (clocks or {}).get('sm_mhz') or 1965.0
    floor_ms = tiles * 3 * rk / 148.0 * clk_tile_rhs / (sm_mhz * 1e3)
    roofline['operand_path'] = {
        'bound': 'tcgen05 shared-memory operand fetch', 'floor_clk_per_tile_rhs': clk_tile_rhs,
        'achieved_clk_per_tile_rhs': kernel_ms * sm_mhz * 1e3 / (tiles * 3 * rk / 148.0),
        'floor_ms': floor_ms, 'frac': floor_ms / kernel_ms, 'sm_mhz': sm_mhz,
        'ncu': 'sm__pipe_tc_cycles_active / l1tex__data_pipe_tc_wavefronts_mem_shared in profiles/r02/tc_*_ncu_full.json'}
  else:
    roofline = dict(hbm)
  roofline.update({'kernel': kernel_name, 'kernel_ms': kernel_ms})
  line = {
      'metric': 'grid-point-steps/sec', 'value': value, 'unit': 'grid-point-steps/s',
      'n_gpus': world, 'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms_per_step,
      'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic', 'config': workload_config(args, batch),
      'clocks': clocks,
      'e2e': {'value': e2e_value, 'unit': 'grid-point-steps/s',
              'h2d_bytes_per_step': batch * n * 4,
              'd2h_bytes_per_step': batch * world * n * 4 + (batch * 4 if world == 1 else 0)},
      'gpu_launches': int(launches),
      'engine': engine,
      'roofline': roofline,
      'roofline_hbm': hbm,
      'compute': {'bound': 'fp32-ffma', 'achieved_tflops': achieved_tf,
                  'peak_tflops': FP32_PEAK_TFLOPS, 'frac': achieved_tf / FP32_PEAK_TFLOPS,
                  'flops_per_grid_point_step': fl, 'peak_kind': 'nominal 148 SM x 128 FMA x 1.965 GHz',
                  'note': 'FP32-equivalent algorithmic FLOPs against the CUDA-core peak (can exceed 1 on the tensor engine)'},
      'launch': shape, 'wall_s': wall, 'final_gather': gather,
  }
  line['parity_check'] = parity_check(args, batch, final_rows, args.steps * rk)
  if world == 1 and not args.no_cpu:
    samples = 16                              # ~10-20 s of single-core work (the bounded sample of the batch)
    _cpu_one_sample((args.workload, 0, 2))     # imports, first-call costs
    rate, elapsed, nfev = cpu_reference_rate(args.workload, min(rk, 500), samples)
    line['cpu_baseline'] = {
        'value': rate, 'unit': 'grid-point-steps/s', 'cores': 1, 'kind': 'port',
        'sample': '%d samples x %d RK3 steps, N=%d, scipy solve_ivp(RK23, max_step=dt) on the oracle port, '
                  '1 process, BLAS limited to 1 thread, %.1f s' % (samples, min(rk, 500), n, elapsed)}
  if world == 1 and args.extra:
    others = {}
    for name in args.extra.split(','):
      try:
        others[name] = quick_rate(name)
      except Exception as e:                   # a secondary workload must not take the headline down
        others[name] = {'error': '%s: %s' % (type(e).__name__, e)}
    line['other_workloads'] = others
  emit(line)


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
  ap.add_argument('--workload', default='c2', choices=sorted(workloads().WORKLOADS))
  ap.add_argument('--rk-steps', type=int, default=500, dest='rk_steps',
                  help='RK3 steps per launch; the default makes --steps 20 the configured 10 000 steps')
  ap.add_argument('--batch', type=int, default=0, help='per-GPU batch override')
  ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                  help='weak: the per-GPU batch is fixed; strong: %d rows in all, split over the GPUs' % STRONG_BATCH)
  ap.add_argument('--engine', default='auto', choices=['auto', 'ffma', 'tensor', 'tensor_f16x2', 'tensor_f16'])
  ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
  ap.add_argument('--extra', default='c3,c4,c5,c1b,c2s',
                  help="other workloads measured briefly at N=1 and reported under 'other_workloads' ('' = none)")
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
