"""Mint tests/golden/long_horizon.npz: the BASELINE.json configurations 2-4 integrated by the ORACLE
over their full configured horizon (10 000 Bogacki-Shampine RK3 steps), for a few rows of the 4096-row
batch the gpu tests and bench.py integrate.

    python tests/golden/make_long_horizon.py          # ~10 minutes on 8 cores

Test infrastructure.  Needs only NumPy and oracle/ (not /root/reference): the oracle itself is pinned
to the reference by make_golden.py / test_oracle_golden.py; this file extends that pin in TIME.
Per case it stores
  <case>/u0      float32 [rows, N]   the initial rows (checked against workloads.horizon_rows)
  <case>/rows    int32   [rows]      their indices in the batch = the forcing seeds
  <case>/f32     float64 [10, rows, N]   float64 state, float32 right-hand side: the reference's
                                         arithmetic (SciPy carries y in float64, the graph is float32)
  <case>/f64     float64 [10, rows, N]   everything in float64: the exact trajectory of the scheme
snapshots every 1000 steps.  |f32 - f64| is the drift the reference's own float32 graph accumulates;
the gpu tests bound the CUDA engines' distance to f64 by it.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = ('c2', 'c3', 'c4', 'c2_unforced')
SAVE_EVERY = 1000


def _case(args):
  name, precision = args
  import ddd1d_b200.workloads as wl
  from oracle import pde_oracle as O
  try:
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=1)
  except ImportError:
    pass
  workload = 'c2' if name == 'c2_unforced' else name
  kind, variant, n, batch, dt, mode = wl.WORKLOADS[workload]
  if name == 'c2_unforced':
    u0, picks = wl.decaying_rows(n), np.arange(4)
  else:
    picks = np.asarray(wl.HORIZON_PICKS)
    u0 = wl.horizon_rows(workload)[picks]
  eqs = [O.EquationSpec(kind, variant, num_points=n, random_seed=int(s)) for s in picks]
  net, weights = O.NetSpec(), wl.synthetic_weights(kind)
  forced = kind == 'burgers' and name != 'c2_unforced'
  real = np.float32 if precision == 'f32' else np.float64

  def rhs(t, y):
    y_t = O.predict_time_derivative(np.asarray(y, dtype=real), eqs[0], net, weights, dtype=real)
    if forced:
      y_t = y_t + np.stack([e.forcing(real(t), dtype=real) for e in eqs])
    return y_t

  t0 = time.time()
  out = O.fixed_step_integrate(rhs, u0, 0.0, dt, wl.FULL_STEPS, SAVE_EVERY)
  return name, precision, u0, picks.astype(np.int32), out, time.time() - t0


def main():
  cases = [c for c in CASES if len(sys.argv) < 2 or c in sys.argv[1:]]      # optional: only the named cases
  jobs = [(c, p) for c in cases for p in ('f32', 'f64')]
  arrays = {}
  target = os.path.join(HERE, 'long_horizon.npz')
  if len(cases) < len(CASES) and os.path.exists(target):
    with np.load(target) as f:
      arrays.update({k: f[k] for k in f.files})
  with mp.get_context('spawn').Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
    for name, precision, u0, picks, out, secs in pool.imap_unordered(_case, jobs):
      arrays['%s/u0' % name] = u0
      arrays['%s/rows' % name] = picks
      arrays['%s/%s' % (name, precision)] = out
      print('%s %s: %d snapshots, max |u| %.3f, %.0f s' % (name, precision, out.shape[0], np.abs(out).max(), secs),
            flush=True)
  for name in cases:
    a, b = arrays['%s/f32' % name], arrays['%s/f64' % name]
    assert np.isfinite(a).all() and np.isfinite(b).all(), name
    print('%s: float32-graph drift over the horizon, relative L-inf per snapshot: %s' % (
        name, ' '.join('%.1e' % (np.abs(x - y).max() / np.abs(y).max()) for x, y in zip(a, b))))
  np.savez_compressed(target, **arrays)


if __name__ == '__main__':
  main()
