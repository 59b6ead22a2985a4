"""Parity over the CONFIGURED horizon: BASELINE.json configs 2-4 (Burgers N=256, KdV N=128, KS N=512;
4096 rows, 10 000 Bogacki-Shampine RK3 steps) on the tensor engine, against the oracle's trajectories in
tests/golden/long_horizon.npz (minted by tests/golden/make_long_horizon.py; float64 state, float32 and
float64 right-hand sides), snapshot by snapshot every 1000 steps.

Rows compared: 0, 1, 2047 and 4095 of the batch.  Row 0 of the Burgers batch starts from the reference's
initial_value() = zeros, so the tensor engine's per-row activation bound (fixed at twice the row maximum
when a row starts) is outgrown again and again while forcing builds the solution up; row 1 of the KS batch
is a mode the u_xxxx term damps by orders of magnitude and the unforced Burgers rows decay under
viscosity, so the periodic "has the row fallen 256x below its bound" re-calibration fires; row 4095 is the
last row of the last wave.

Tolerance.  integrate.py carries a float64 state through a float32 graph; two correct float32 evaluations
with different summation order drift apart by about as much as either drifts from the exact (float64)
trajectory.  That drift of the reference's own arithmetic is in the fixture (|f32 - f64| relative to the
row's amplitude: up to 4.4e-5 for C2, 1e-7 for C3, 1.6e-4 for the O(1) rows of C4 and 2.8e-3 for its
damped row over the 10 000 steps), so the assertions are, per row and per snapshot:
  (1) the CUDA trajectory is at most FACTOR x as far from the float64 trajectory as the float32 oracle
      is (floor FLOOR);
  (2) it is within ABS_TOL of the float32 oracle: BASELINE.md section 4's 1e-4 where the float32 graph's
      own drift leaves room for it (C3; the unforced rows), 2e-4 for C2 (twice the drift there is already
      0.9e-4) and 5e-4 for C4 -- except on rows where 4 x the drift exceeds it.
"""
import os

import numpy as np
import pytest

from tests import gpu_helpers as G

pytestmark = pytest.mark.gpu

FACTOR = 3.0
FLOOR = 5e-6
ABS_TOL = {'c2': 2e-4, 'c3': 1e-4, 'c4': 5e-4, 'c2_unforced': 1e-4}     # vs the float32 oracle, relative L-inf
SAVE_EVERY = 1000


@pytest.fixture(scope='module')
def horizon(golden):
  return golden('long_horizon')


def _solver(workload, engine, rows, forcing=True):
  import ddd1d_b200.workloads as wl
  from ddd1d_b200 import runtime
  kind, variant, n, _, dt, _ = wl.WORKLOADS[workload]
  eqs = [G.product_equation(kind, variant, n, seed=int(s)) for s in rows]
  solver = runtime.learned_solver(eqs, G.product_hparams(kind, variant, n), wl.synthetic_weights(kind),
                                  engine=engine, forcing=forcing)
  assert solver.engine() == engine
  return solver, dt


def _check(case, got, want32, want64, what):
  """got, want*: [snapshots, rows, N]."""
  worst = 0.0
  for i in range(got.shape[0]):
    for r in range(got.shape[1]):        # every row against its OWN amplitude: decayed rows count like O(1) rows
      scale = np.abs(want64[i, r]).max()
      assert scale > 1e-30, 'row %d left the float32 range' % r
      e_got = np.abs(got[i, r] - want64[i, r]).max() / scale
      e_ref = np.abs(want32[i, r] - want64[i, r]).max() / scale
      e_abs = np.abs(got[i, r] - want32[i, r]).max() / scale
      worst = max(worst, e_abs)
      where = '%s row %d step %d' % (what, r, (i + 1) * SAVE_EVERY)
      assert e_got <= max(FACTOR * e_ref, FLOOR), (
          '%s: %.2e from the float64 trajectory, the float32 oracle is %.2e' % (where, e_got, e_ref))
      assert e_abs <= max(ABS_TOL[case], 4 * e_ref), '%s: %.2e from the float32 oracle' % (where, e_abs)
  print('%s: worst relative L-inf (per row) vs the float32 oracle over 10 snapshots %.2e' % (what, worst))


@pytest.mark.parametrize('workload', ('c2', 'c3', 'c4'))
def test_configured_horizon_tensor_engine(workload, horizon):
  """The full batch of the configuration on the tensor engine, 10 000 steps in one launch."""
  import ddd1d_b200.workloads as wl
  rows = wl.horizon_rows(workload)
  picks = horizon['%s/rows' % workload]
  np.testing.assert_array_equal(rows[picks], horizon['%s/u0' % workload])
  solver, dt = _solver(workload, 'tensor', range(rows.shape[0]))
  snaps, bad = solver.integrate(rows, 0.0, dt, wl.FULL_STEPS, SAVE_EVERY, return_first_bad=True)
  assert (bad.cpu().numpy() == -1).all()
  got = snaps[:, picks.tolist()].cpu().numpy().astype(np.float64)
  _check(workload, got, horizon['%s/f32' % workload], horizon['%s/f64' % workload], workload + ' tensor')
  # the same trajectory launch by launch (what bench.py times): state carried through float32 snapshots
  state, t = rows, 0.0
  for i in range(4):
    state = solver.integrate(state, t, dt, 500, 500)[0]
    t += 500 * dt
  two = state[picks.tolist()].cpu().numpy().astype(np.float64)
  scale = np.abs(horizon['%s/f64' % workload][1]).max()
  assert np.abs(two - got[1]).max() / scale < 2e-6     # float32 rounding of the carried state, 4 times
  solver.close()


@pytest.mark.parametrize('workload', ('c2', 'c3', 'c4'))
def test_configured_horizon_ffma_engine(workload, horizon):
  """The FP32-FFMA engine on the compared rows alone (rows are independent; their forcing seeds are
  their batch indices)."""
  import ddd1d_b200.workloads as wl
  picks = horizon['%s/rows' % workload]
  solver, dt = _solver(workload, 'ffma', picks)
  snaps = solver.integrate(horizon['%s/u0' % workload], 0.0, dt, wl.FULL_STEPS, SAVE_EVERY)
  _check(workload, snaps.cpu().numpy().astype(np.float64), horizon['%s/f32' % workload],
         horizon['%s/f64' % workload], workload + ' ffma')
  solver.close()


@pytest.mark.parametrize('engine', ('tensor', 'ffma'))
def test_decaying_rows(engine, horizon):
  """Unforced Burgers rows that viscosity damps by many orders of magnitude: the activation bound has to
  follow the row down (relative accuracy must hold at 1e-10 amplitudes as it does at 1)."""
  import ddd1d_b200.workloads as wl
  u0 = wl.decaying_rows()
  np.testing.assert_array_equal(u0, horizon['c2_unforced/u0'])
  solver, dt = _solver('c2', engine, range(4), forcing=False)
  snaps, bad = solver.integrate(u0, 0.0, dt, wl.FULL_STEPS, SAVE_EVERY, return_first_bad=True)
  assert (bad.cpu().numpy() == -1).all()
  _check('c2_unforced', snaps.cpu().numpy().astype(np.float64), horizon['c2_unforced/f32'],
         horizon['c2_unforced/f64'], 'unforced Burgers ' + engine)
  solver.close()
