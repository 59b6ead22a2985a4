"""Margins of the long-horizon parity test (tests/test_gpu_long_horizon.py) as a table: per workload the worst
ratio e_got / max(3 e_ref, 5e-6) over rows and snapshots (must stay <= 1), plus the worst errors.  Used to
compare library variants.  python scripts/horizon_table.py [c2 c3 c4]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from tests import test_gpu_long_horizon as T
import ddd1d_b200.workloads as wl

f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'long_horizon.npz'))
for w in (sys.argv[1:] or ['c2', 'c3', 'c4']):
  rows = wl.horizon_rows(w)
  picks = f['%s/rows' % w]
  solver, dt = T._solver(w, os.environ.get('ENGINE', 'tensor'), range(rows.shape[0]))
  snaps = solver.integrate(rows, 0.0, dt, wl.FULL_STEPS, T.SAVE_EVERY)
  got = snaps[:, picks.tolist()].cpu().numpy().astype(np.float64)
  want32, want64 = f['%s/f32' % w], f['%s/f64' % w]
  worst = (0, None)
  eg_max = er_max = ea_max = 0.0
  for i in range(got.shape[0]):
    for r in range(got.shape[1]):
      scale = np.abs(want64[i, r]).max()
      e_got = np.abs(got[i, r] - want64[i, r]).max() / scale
      e_ref = np.abs(want32[i, r] - want64[i, r]).max() / scale
      e_abs = np.abs(got[i, r] - want32[i, r]).max() / scale
      ratio = e_got / max(T.FACTOR * e_ref, T.FLOOR)
      if ratio > worst[0]: worst = (ratio, (i, r, e_got, e_ref))
      eg_max, er_max, ea_max = max(eg_max, e_got), max(er_max, e_ref), max(ea_max, e_abs)
  print('%s: worst margin ratio %.3f at snapshot %d row %d (e_got %.2e, e_ref %.2e); max e_got %.2e  max e_ref %.2e  max |got - f32| %.2e'
        % ((w, worst[0]) + worst[1] + (eg_max, er_max, ea_max)), flush=True)
  solver.close()
