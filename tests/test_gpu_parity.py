"""Parity of the CUDA path (through the C ABI, via ddd1d_b200.runtime.RowSolver ->
ctypes -> libddd1d.so) against (a) fixtures minted from the reference's own code
and (b) the NumPy oracle on the same seeded inputs.

Tolerances (float32 arithmetic; the reference's TF graph is float32 too):
  RHS_TOL   1e-5  relative L-inf of dy/dt, coefficients and derivatives per call
            (BASELINE.md section 4);
  TRAJ_TOL  1e-4  relative L-inf of a fixed-step trajectory vs the oracle's
            fixed-step trajectory (float64 state on both sides).
"""
import numpy as np
import pytest

from oracle import pde_oracle as O
from tests.helpers import KINDS, VARIANTS, assert_f32_faithful, net_from_json, rel_err, weights_from
from tests import gpu_helpers as G

pytestmark = pytest.mark.gpu

RHS_TOL = 1e-5
FLUX_TOL = 3e-5   # dy/dt of flux-difference forms: float32 rounding of the flux is divided by dx
TRAJ_TOL = 1e-4


def cpu(t):
  return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------
# per-call parity: coefficients, derivatives, dy/dt
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('variant', VARIANTS)
@pytest.mark.parametrize('n', (32, 64))
@pytest.mark.parametrize('engine', ('ffma', 'tensor'))
def test_learned_against_reference_fixture(golden, monkeypatch, kind, variant, n, engine):
  """Both engines at the reference's own grid sizes; the tensor engine packs 4 (N=32) or 2 (N=64)
  rows into one 128-position MMA tile."""
  from ddd1d_b200 import model, integrate
  monkeypatch.setenv('DDD1D_ENGINE', engine)
  g = golden('learned')
  key = 'default/%s/%s/%d' % (kind, variant, n)
  hp = G.product_hparams(kind, variant, n)
  w = weights_from(g, key)
  u = g[key + '/u']
  assert rel_err(cpu(model.predict_coefficients(u, hp, w)), g[key + '/coefficients']) < RHS_TOL
  assert rel_err(cpu(model.predict_space_derivatives(u, hp, w)), g[key + '/space_derivatives']) < RHS_TOL
  assert rel_err(cpu(model.predict_time_derivative(u, hp, w)), g[key + '/time_derivative']) < RHS_TOL
  # the Differentiator surface SciPy calls (adds forcing for Burgers)
  eq = G.product_equation(kind, variant, n, seed=7)
  d = integrate.SavedModelDifferentiator(w, eq, hp)
  assert d.solver.engine() == engine
  got = d(float(g[key + '/t']), u[0].astype(np.float64))
  assert got.dtype == np.float64
  assert rel_err(got, g[key + '/differentiator']) < RHS_TOL


@pytest.mark.parametrize('kind,variant', (('ks', 'conservative'), ('ks', 'plain'), ('burgers', 'plain'),
                                          ('burgers', 'conservative'), ('kdv', 'godunov')))
@pytest.mark.parametrize('n', (32, 64))
def test_coefficient_grid_min_size_9_against_reference_fixture(golden, kind, variant, n):
  """hparams.coefficient_grid_min_size = 9 (model.py:445-448; the reference's training_test.py:56 runs it on the
  default conservative KS): 9 centred or 10 staggered points, the 11-slot window of the FFMA engine."""
  from ddd1d_b200 import model, integrate
  g = golden('grid9')
  key = '%s/%s/%d' % (kind, variant, n)
  hp = G.product_hparams(kind, variant, n, coefficient_grid_min_size=9)
  w = weights_from(g, key)
  u = g[key + '/u']
  coefs = cpu(model.predict_coefficients(u, hp, w))
  assert coefs.shape == g[key + '/coefficients'].shape
  assert rel_err(coefs, g[key + '/coefficients']) < RHS_TOL
  assert rel_err(cpu(model.predict_space_derivatives(u, hp, w)), g[key + '/space_derivatives']) < RHS_TOL
  assert rel_err(cpu(model.predict_time_derivative(u, hp, w)), g[key + '/time_derivative']) < FLUX_TOL
  eq = G.product_equation(kind, variant, n, seed=9)
  d = integrate.SavedModelDifferentiator(w, eq, hp)
  assert d.solver.engine() == 'ffma'
  assert rel_err(d(0.23, u[0].astype(np.float64)), g[key + '/differentiator']) < FLUX_TOL
  # and a short fused integration against the oracle
  oeq = G.oracle_equation(kind, variant, n, seed=9)
  net = O.NetSpec(coefficient_grid_min_size=9)
  dt = {'burgers': 1e-3, 'kdv': 2.5e-5, 'ks': 1e-5}[kind]
  w_small = [(k, b) for k, b in w[:-1]] + [(w[-1][0] * 0.1, w[-1][1] * 0.1)]
  solver = integrate.BatchIntegrator.learned([eq], hp, w_small)
  got = cpu(solver.integrate(u[:1], 0.0, dt, 6, 3))
  want = O.fixed_step_integrate(O.batched_rhs([oeq], net, w_small, mode='learned'), u[:1], 0.0, dt, 6, 3)
  assert rel_err(got, want) < TRAJ_TOL


def test_learned_hparam_variants_against_reference_fixture(golden):
  from ddd1d_b200 import model
  g = golden('learned')
  names = sorted({k.split('/')[0] for k in g.files} - {'default'})
  for name in names:
    for variant in ('plain', 'conservative'):
      key = '%s/burgers/%s/32' % (name, variant)
      if key + '/u' not in g.files:
        continue
      overrides = dict(__import__('json').loads(str(g[key + '/hparams'])))
      hp = G.product_hparams('burgers', variant, 32, **overrides)
      w = weights_from(g, key)
      u = g[key + '/u']
      assert rel_err(cpu(model.predict_coefficients(u, hp, w)), g[key + '/coefficients']) < RHS_TOL, key
      assert rel_err(cpu(model.predict_time_derivative(u, hp, w)), g[key + '/time_derivative']) < RHS_TOL, key


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('variant', VARIANTS)
def test_baseline_against_reference_fixture(golden, kind, variant):
  from ddd1d_b200 import model, integrate
  g = golden('baseline')
  key = '%s/%s/32' % (kind, variant)
  eq = G.product_equation(kind, variant, 32, seed=11)
  u = g[key + '/u']
  for acc in (1, 3):
    sd = model.baseline_space_derivatives(u, eq, acc)
    assert rel_err(cpu(sd), g['%s/acc%d/space_derivatives' % (key, acc)]) < RHS_TOL
    d = integrate.PolynomialDifferentiator(eq, acc)
    assert rel_err(d(1.25, u[0].astype(np.float64)), g['%s/acc%d/differentiator' % (key, acc)]) < RHS_TOL
    named = d.calculate_space_derivatives(u[0])
    assert sorted(named) == sorted(eq.DERIVATIVE_NAMES)


def test_weno_against_reference_fixture(golden):
  from ddd1d_b200 import weno, model, integrate, equations
  g = golden('pointwise')
  u = g['weno/u']
  left, right = weno.reconstruct_left(u), weno.reconstruct_right(u)       # float64 kernel
  np.testing.assert_allclose(left, g['weno/left'], rtol=1e-12, atol=1e-12)
  np.testing.assert_allclose(right, g['weno/right'], rtol=1e-12, atol=1e-12)
  l32, r32 = weno.reconstruct_both(u.astype(np.float32))                  # float32 kernel
  np.testing.assert_allclose(l32, g['weno/left_f32'], rtol=0, atol=3e-5)
  np.testing.assert_allclose(r32, g['weno/right_f32'], rtol=0, atol=3e-5)
  # the "exact" dispatch of baseline_space_derivatives (WENO in float32, model.py:81-97)
  gb = golden('baseline')
  eq = equations.GodunovBurgersEquation(32, random_seed=11)
  sd = model.baseline_space_derivatives(gb['burgers/godunov/32/exact/u'], eq, None)
  assert rel_err(cpu(sd), gb['burgers/godunov/32/exact/space_derivatives']) < 2e-5
  # WENODifferentiator (integrate.py:124-140): float64 reconstruction / flux / forcing with the
  # float32 4-point stencil, as in the reference; and the all-float32 kernel at a looser tolerance
  gt = golden('trajectories')
  eqw = equations.GodunovBurgersEquation(64, random_seed=1)
  d = integrate.WENODifferentiator(eqw)
  # (the float32 4-point u_x stencil, a float32 TF graph in the reference too, sets the floor here)
  assert rel_err(d(0.4, gt['weno_burgers/rhs_u']), gt['weno_burgers/rhs']) < 3e-6
  d32 = integrate.WENODifferentiator(eqw, weno_real='float32')
  assert rel_err(d32(0.4, gt['weno_burgers/rhs_u']), gt['weno_burgers/rhs']) < 2e-5
  # the reference's own consistency test (integrate_test.py:146-155): exact == weno
  y, nfev = integrate.odeint(eqw.initial_value(), d, gt['weno_burgers/times'])
  assert nfev == int(gt['weno_burgers/nfev'])
  np.testing.assert_allclose(y, gt['weno_burgers/y'], rtol=0, atol=1e-6)
  # Godunov KdV / KS through the same float64 path; their float32 u_xx / u_xxx stencils (1/dx^n) set
  # a higher floor for how closely two float32 evaluations can agree
  for kind, tol in (('kdv', 1e-5), ('ks', 5e-5)):
    eqk = equations.FLUX_EQUATION_TYPES[kind](64, random_seed=2)
    oeq = G.oracle_equation(kind, 'godunov', 64, seed=2)
    u = G.smooth_rows(1, 64, seed=3)[0].astype(np.float64)
    assert rel_err(integrate.WENODifferentiator(eqk)(0.1, u), O.WENODifferentiator(oeq)(0.1, u)) < tol


# ---------------------------------------------------------------------------------
# trajectories
# ---------------------------------------------------------------------------------
def _fixed_step_case(kind, variant, n, batch, steps, dt, mode, seed=0, scheme='rk3', tol=TRAJ_TOL):
  from ddd1d_b200 import integrate
  eqs = [G.product_equation(kind, variant, n, seed=s) for s in range(batch)]
  oeqs = [G.oracle_equation(kind, variant, n, seed=s) for s in range(batch)]
  u0 = G.smooth_rows(batch, n, seed=seed)
  if mode == 'learned':
    net = O.NetSpec()
    w = O.glorot_weights(oeqs[0], net, seed=seed + 1, last_layer_scale=0.01)
    solver = integrate.BatchIntegrator.learned(eqs, G.product_hparams(kind, variant, n), w)
    rhs = O.batched_rhs(oeqs, net, w, mode='learned')
  elif mode == 'fd':
    solver = integrate.BatchIntegrator.baseline(eqs, 1)
    rhs = O.batched_rhs(oeqs, mode='fd', accuracy_order=1)
  else:
    solver = integrate.BatchIntegrator.weno(eqs)
    rhs = O.batched_rhs(oeqs, mode='weno', weno_dtype=np.float32)
  save_every = max(1, steps // 2)
  got, bad = solver.integrate(u0, 0.1, dt, steps, save_every, scheme, return_first_bad=True)
  want = O.fixed_step_integrate(rhs, u0, 0.1, dt, steps, save_every, scheme=scheme)
  assert got.shape == want.shape
  assert (cpu(bad) == -1).all()
  err = rel_err(cpu(got), want)
  assert err < tol, (kind, variant, mode, err)
  return err


@pytest.mark.parametrize('kind,dt', (('burgers', 1e-3), ('kdv', 2.5e-5), ('ks', 1e-5)))
@pytest.mark.parametrize('variant', VARIANTS)
def test_fixed_step_learned(kind, dt, variant):
  _fixed_step_case(kind, variant, 64, 5, 40, dt, 'learned')


@pytest.mark.parametrize('kind,dt', (('burgers', 1e-3), ('kdv', 2.5e-5), ('ks', 1e-5)))
@pytest.mark.parametrize('variant', VARIANTS)
def test_fixed_step_baseline(kind, dt, variant):
  _fixed_step_case(kind, variant, 64, 5, 60, dt, 'fd')


@pytest.mark.parametrize('kind,dt', (('burgers', 1e-3), ('kdv', 2.5e-5), ('ks', 1e-5)))
def test_fixed_step_weno(kind, dt):
  _fixed_step_case(kind, 'godunov', 64, 5, 60, dt, 'weno', tol=2e-4)


@pytest.mark.parametrize('scheme', ('midpoint', 'euler', 'rk4'))
def test_other_schemes(scheme):
  _fixed_step_case('burgers', 'plain', 32, 3, 30, 1e-3, 'learned', scheme=scheme)


def test_scipy_driven_c1_matches_reference_fixture(golden):
  """BASELINE config 1: Burgers FD accuracy 1, N=64, T=2 through the reference's
  own driver loop (SciPy RK23, max_step=0.01) with the GPU Differentiator."""
  from ddd1d_b200 import integrate, equations
  g = golden('trajectories')
  for tag, seed in (('c1', 0), ('c1_seed2', 2)):
    eq = equations.BurgersEquation(64, random_seed=seed)
    ds = integrate.integrate_baseline(eq, times=g['c1/times'])
    y = np.asarray(ds['y'].data)
    assert int(np.asarray(ds['num_evals'].data if hasattr(ds['num_evals'], 'data') else ds['num_evals'])) == int(g[tag + '/nfev'])
    np.testing.assert_allclose(y, g[tag + '/y'], rtol=0, atol=5e-5)
    assert abs(y.mean(axis=1)).max() < 1e-3          # integrate_test.py:182-185


def test_scipy_driven_learned_matches_reference_fixture(golden):
  from ddd1d_b200 import integrate, equations
  g = golden('trajectories')
  eq = equations.BurgersEquation(32, random_seed=4)
  w = weights_from(g, 'learned_burgers')
  d = integrate.SavedModelDifferentiator(w, eq, G.product_hparams('burgers', 'plain', 32))
  y, nfev = integrate.odeint(eq.initial_value(), d, g['learned_burgers/times'])
  assert nfev == int(g['learned_burgers/nfev'])
  np.testing.assert_allclose(y, g['learned_burgers/y'], rtol=0, atol=5e-5)
  eq = equations.KdVEquation(32, random_seed=2)
  w = weights_from(g, 'learned_kdv')
  d = integrate.SavedModelDifferentiator(w, eq, G.product_hparams('kdv', 'plain', 32))
  y, nfev = integrate.odeint(eq.initial_value(), d, g['learned_kdv/times'])
  assert nfev == int(g['learned_kdv/nfev'])
  np.testing.assert_allclose(y, g['learned_kdv/y'], rtol=0, atol=1e-4)


def test_fixed_step_c1_twin_matches_scipy_fixture(golden):
  """The fused fixed-step kernel (dt=0.01, 200 steps) lands on the reference's adaptive
  result for C1 within SciPy's own tolerance (rtol=1e-3): the controller sits at max_step."""
  from ddd1d_b200 import integrate, equations
  g = golden('trajectories')
  eqs = [equations.BurgersEquation(64, random_seed=s) for s in (0, 2)]
  ds = integrate.BatchIntegrator.baseline(eqs, 1).integrate_times(times=g['c1/times'], dt=0.01)
  y = np.asarray(ds['y'].data)
  assert y.shape == (2, 5, 64)
  np.testing.assert_allclose(y[0], g['c1/y'], rtol=0, atol=5e-4)
  np.testing.assert_allclose(y[1], g['c1_seed2/y'], rtol=0, atol=5e-4)


def test_device_adaptive_rk23_matches_scipy_fixtures(golden):
  """ddd1d_integrate_adaptive: SciPy's RK23 controller, FSAL and dense output re-implemented
  per row on the device.  Against trajectories the reference's own odeint produced (with
  the reference's graph code on the NumPy shim): same number of RHS evaluations, same
  samples.  c1_seed1 (nfev 1031) is a run where the controller is NOT pinned at max_step."""
  from ddd1d_b200 import integrate, equations
  g = golden('trajectories')
  eqs = [equations.BurgersEquation(64, random_seed=s) for s in (0, 1, 2)]
  y, nfev = integrate.BatchIntegrator.baseline(eqs, 1).odeint(times=g['c1/times'])
  assert y.shape == (3, 5, 64) and y.dtype == np.float64
  for i, tag in enumerate(('c1', 'c1_seed1', 'c1_seed2')):
    assert nfev[i] == int(g[tag + '/nfev']), (tag, nfev[i])
    np.testing.assert_allclose(y[i], g[tag + '/y'], rtol=0, atol=5e-5)
  eq = equations.ConservativeBurgersEquation(32, resample_factor=4, random_seed=3)
  y, nfev = integrate.BatchIntegrator.baseline([eq], 1).odeint(times=g['cons_burgers/times'])
  assert nfev[0] == int(g['cons_burgers/nfev'])
  np.testing.assert_allclose(y[0], g['cons_burgers/y'], rtol=0, atol=5e-5)
  eq = equations.BurgersEquation(32, random_seed=4)
  solver = integrate.BatchIntegrator.learned([eq], G.product_hparams('burgers', 'plain', 32),
                                             weights_from(g, 'learned_burgers'))
  y, nfev = solver.odeint(times=g['learned_burgers/times'])
  assert nfev[0] == int(g['learned_burgers/nfev'])
  np.testing.assert_allclose(y[0], g['learned_burgers/y'], rtol=0, atol=5e-5)
  eq = equations.KdVEquation(32, random_seed=2)
  solver = integrate.BatchIntegrator.learned([eq], G.product_hparams('kdv', 'plain', 32),
                                             weights_from(g, 'learned_kdv'))
  y, nfev = solver.odeint(times=g['learned_kdv/times'])
  assert nfev[0] == int(g['learned_kdv/nfev'])
  np.testing.assert_allclose(y[0], g['learned_kdv/y'], rtol=0, atol=1e-4)
  eq = equations.GodunovBurgersEquation(64, random_seed=1)
  y, nfev = integrate.BatchIntegrator.weno([eq]).odeint(times=g['weno_burgers/times'])
  assert nfev[0] == int(g['weno_burgers/nfev'])
  np.testing.assert_allclose(y[0], g['weno_burgers/y'], rtol=0, atol=5e-5)   # float32 WENO vs float64 reference
  y, nfev = integrate.BatchIntegrator.weno([eq], weno_real='float64').odeint(times=g['weno_burgers/times'])
  assert nfev[0] == int(g['weno_burgers/nfev'])
  np.testing.assert_allclose(y[0], g['weno_burgers/y'], rtol=0, atol=1e-6)   # float64 WENO twin


def test_device_adaptive_nan_padding_and_status():
  """A run whose step size underflows stops early: later samples are NaN, status -1
  (integrate.py:161-167), other rows of the batch are unaffected."""
  from ddd1d_b200 import integrate, equations
  eqs = [equations.KdVEquation(64, random_seed=s) for s in range(2)]
  solver = integrate.BatchIntegrator.baseline(eqs, 1)
  u0 = solver.initial_values().astype(np.float64)
  u0[1, 10] = np.inf                      # poisons row 1 only
  times = np.linspace(0, 0.02, 5)
  y, nfev, status = solver.solver.odeint(u0, times)
  y, status = cpu(y), cpu(status)
  assert status[0] == 0 and np.isfinite(y[:, 0]).all()
  assert status[1] == -1 and np.isnan(y[1:, 1]).all()
  want, nf = O.odeint(u0[0], O.PolynomialDifferentiator(G.oracle_equation('kdv', 'plain', 64, seed=0), 1), times)
  assert int(cpu(nfev)[0]) == nf
  np.testing.assert_allclose(y[:, 0], want, rtol=0, atol=5e-5)


# ---------------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties + oracle on a row subset
# ---------------------------------------------------------------------------------
def _c2_solver(batch, n=256):
  from ddd1d_b200 import integrate
  eqs = [G.product_equation('burgers', 'plain', n, seed=s) for s in range(batch)]
  oeq = G.oracle_equation('burgers', 'plain', n)
  w = O.glorot_weights(oeq, O.NetSpec(), seed=0, last_layer_scale=0.01)
  return integrate.BatchIntegrator.learned(eqs, G.product_hparams('burgers', 'plain', n), w), w


def test_c2_full_batch_properties():
  """BASELINE config 2 shape (Burgers learned, N=256, batch=4096), short run."""
  import torch
  batch, n, steps = 4096, 256, 6
  solver, w = _c2_solver(batch, n)
  u0 = G.smooth_rows(batch, n, seed=5)
  full = solver.integrate(u0, 0.0, 1e-3, steps, steps)[0]
  assert torch.isfinite(full).all()
  # (1) a row's trajectory does not depend on the rest of the batch or on the CTA it lands on
  pick = [0, 1, 147, 148, 149, 2047, 4095]
  for i in pick:
    alone = solver.integrate(u0[i:i + 1], 0.0, 1e-3, steps, steps, sample_offset=i)[0, 0]
    assert torch.equal(alone, full[i]), i
  # (2) oracle on the same subset
  oeqs = [G.oracle_equation('burgers', 'plain', n, seed=i) for i in pick]
  rhs = O.batched_rhs(oeqs, O.NetSpec(), w, mode='learned')
  want = O.fixed_step_integrate(rhs, u0[pick], 0.0, 1e-3, steps, steps)[0]
  assert rel_err(cpu(full[pick]), want) < TRAJ_TOL
  # (3) determinism
  again = solver.integrate(u0, 0.0, 1e-3, steps, steps)[0]
  assert torch.equal(again, full)


def test_coefficients_satisfy_accuracy_constraints_at_full_size():
  """polynomials_test.py:94-104 lifted to the network output: A @ coef == b at every
  grid point of a [4096, 256] batch."""
  from ddd1d_b200 import model, polynomials, runtime
  n, batch = 256, 4096
  hp = G.product_hparams('burgers', 'plain', n)
  eq = G.product_equation('burgers', 'plain', n)
  oeq = G.oracle_equation('burgers', 'plain', n)
  w = O.glorot_weights(oeq, O.NetSpec(), seed=3, last_layer_scale=0.1)
  coefs = model.predict_coefficients(G.smooth_rows(batch, n, seed=9), hp, w)
  assert tuple(coefs.shape) == (batch, n, 2, 7)
  grid = runtime.coefficient_grid(eq, hp)
  import torch
  for d, order in enumerate(eq.DERIVATIVE_ORDERS):
    a, b = polynomials.constraints(grid, polynomials.Method.FINITE_DIFFERENCES, order, 1)
    a_t = torch.as_tensor(a, device=coefs.device, dtype=torch.float64)
    resid = coefs[:, :, d, :].double() @ a_t.T - torch.as_tensor(b, device=coefs.device)
    scale = (coefs[:, :, d, :].double().abs() @ a_t.abs().T).max()
    assert float(resid.abs().max() / scale) < 1e-5


def test_translation_equivariance_and_conservation():
  """Unforced equations commute with periodic shifts; conservative forms keep sum(u)."""
  import torch
  from ddd1d_b200 import integrate
  n, batch = 128, 64
  for kind, dt in (('kdv', 2.5e-5), ('ks', 1e-5)):
    for variant in VARIANTS:
      eqs = [G.product_equation(kind, variant, n)]
      oeq = G.oracle_equation(kind, variant, n)
      w = O.glorot_weights(oeq, O.NetSpec(), seed=1, last_layer_scale=0.01)
      solver = integrate.BatchIntegrator.learned(eqs, G.product_hparams(kind, variant, n), w)
      u0 = torch.as_tensor(G.smooth_rows(batch, n, seed=2)).cuda()
      y = solver.integrate(u0, 0.0, dt, 10, 10)[0]
      y_shift = solver.integrate(torch.roll(u0, 17, dims=1), 0.0, dt, 10, 10)[0]
      assert torch.equal(torch.roll(y, 17, dims=1), y_shift), (kind, variant)
      if variant != 'plain':
        drift = (y.double().sum(dim=1) - u0.double().sum(dim=1)).abs().max()
        assert float(drift) < 5e-4 * n, (kind, variant, float(drift))


@pytest.mark.parametrize('kind,n,dt', (('kdv', 128, 2.5e-5), ('ks', 512, 1e-5)))
def test_c3_c4_shapes_against_oracle(kind, n, dt):
  """BASELINE configs 3 and 4 (KdV N=128, KS N=512): per-GPU shard shape, oracle on a subset."""
  import torch
  from ddd1d_b200 import integrate
  batch, steps = 4096, 4
  eqs = [G.product_equation(kind, 'plain', n)]
  oeq = G.oracle_equation(kind, 'plain', n)
  w = O.glorot_weights(oeq, O.NetSpec(), seed=2, last_layer_scale=0.01)
  solver = integrate.BatchIntegrator.learned(eqs, G.product_hparams(kind, 'plain', n), w)
  rs = np.random.RandomState(0)
  u0 = np.stack([O.EquationSpec(kind, num_points=n, random_seed=int(s)).initial_value()
                 for s in rs.randint(0, 1000, size=8)]).astype(np.float32)
  u0 = np.tile(u0, (batch // 8, 1))
  got = solver.integrate(u0, 0.0, dt, steps, steps)[0]
  assert torch.isfinite(got).all()
  rhs = O.batched_rhs([oeq], O.NetSpec(), w, mode='learned')
  want = O.fixed_step_integrate(rhs, u0[:8], 0.0, dt, steps, steps)[0]
  assert rel_err(cpu(got[:8]), want) < TRAJ_TOL
  assert torch.equal(got[:8], got[batch - 8:])


def test_c5_weno_shape_against_oracle():
  """BASELINE config 5 row shape (WENO5 Godunov Burgers, N=2048), reduced batch."""
  import torch
  from ddd1d_b200 import integrate
  n, batch, steps, dt = 2048, 512, 6, 1e-4
  eqs = [G.product_equation('burgers', 'godunov', n, seed=s) for s in range(batch)]
  solver = integrate.BatchIntegrator.weno(eqs)
  u0 = G.smooth_rows(batch, n, seed=4)
  got = solver.integrate(u0, 0.0, dt, steps, steps)[0]
  assert torch.isfinite(got).all()
  pick = [0, 255, 511]
  oeqs = [G.oracle_equation('burgers', 'godunov', n, seed=i) for i in pick]
  rhs = O.batched_rhs(oeqs, mode='weno', weno_dtype=np.float32)
  want = O.fixed_step_integrate(rhs, u0[pick], 0.0, dt, steps, steps)[0]
  assert rel_err(cpu(got[pick]), want) < 2e-4


@pytest.mark.parametrize('n', (384, 512, 1024, 2048))
def test_weno_block_kernel(monkeypatch, n):
  """The register-resident CTA-per-row WENO5 integrator (csrc/ddd1d_weno.cuh; rows of 4 points per thread):
  every Godunov equation against the oracle, every scheme, snapshots, odd batch sizes and divergence
  reporting against the shared-memory CTA-per-row kernel it replaces.  The two differ only in the
  rounding of the WENO weights (reciprocal-and-multiply here, IEEE division there)."""
  from ddd1d_b200 import integrate
  # explicit steps inside the stability limits of the finer grids (dt ~ dx^3 for KdV, dx^4 for KS)
  for kind, dt in (('burgers', 1e-4), ('kdv', 2.5e-5 * min(1.0, (256.0 / n) ** 3)), ('ks', 1e-5 * min(1.0, (512.0 / n) ** 4))):
    _fixed_step_case(kind, 'godunov', n, 3, 12, dt, 'weno', tol=2e-4)
  batch = 301
  eqs = [G.product_equation('burgers', 'godunov', n, seed=s) for s in range(batch)]
  solver = integrate.BatchIntegrator.weno(eqs)
  u0 = G.smooth_rows(batch, n, seed=9)
  u0[7] *= 1e20                                         # this row blows up: divergence is data
  for scheme in ('rk3', 'midpoint', 'euler', 'rk4'):
    monkeypatch.delenv('DDD1D_NO_WENO_BLOCK', raising=False)
    assert solver.solver.launch_shape(batch)['block'] == n // 4
    a, bad_a = solver.integrate(u0, 0.3, 1e-4, 12, 4, scheme, return_first_bad=True)
    monkeypatch.setenv('DDD1D_NO_WENO_BLOCK', '1')
    assert solver.solver.launch_shape(batch)['block'] != n // 4 or n == 2048
    b, bad_b = solver.integrate(u0, 0.3, 1e-4, 12, 4, scheme, return_first_bad=True)
    monkeypatch.delenv('DDD1D_NO_WENO_BLOCK', raising=False)
    a, b = cpu(a), cpu(b)
    np.testing.assert_array_equal(cpu(bad_a), cpu(bad_b))
    ok = np.isfinite(b).all(axis=(0, 2))
    assert ok.sum() == batch - 1 and not ok[7]
    assert rel_err(a[:, ok], b[:, ok]) < 5e-6, scheme
    np.testing.assert_array_equal(np.isfinite(a), np.isfinite(b))


def test_float32_state_carry():
  """model.integrate_ode (model.py:138-159) unrolls the midpoint rule on a float32 tensor: the solution is rounded
  to float32 after every step.  DDD1D_STATE_F32 reproduces that carry (predict_time_evolution /
  baseline_time_evolution use it); the default float64 carry is SciPy's (integrate.py:154).  Over 400 steps the two
  differ measurably, and each matches the oracle run with the same state dtype."""
  from ddd1d_b200 import integrate
  n, steps = 64, 400
  eqs = [G.product_equation('burgers', 'plain', n, seed=s) for s in range(3)]
  oeqs = [G.oracle_equation('burgers', 'plain', n, seed=s) for s in range(3)]
  u0 = G.smooth_rows(3, n, seed=2)
  rhs = O.batched_rhs(oeqs, mode='fd', accuracy_order=1)
  for engine_solver in (integrate.BatchIntegrator.baseline(eqs, 1),):
    got32 = cpu(engine_solver.solver.integrate(u0, 0.0, 1e-3, steps, steps, 'midpoint', float32_state=True))
    got64 = cpu(engine_solver.solver.integrate(u0, 0.0, 1e-3, steps, steps, 'midpoint'))
    want32 = O.fixed_step_integrate(rhs, u0, 0.0, 1e-3, steps, steps, scheme='midpoint', state_dtype=np.float32)
    want64 = O.fixed_step_integrate(rhs, u0, 0.0, 1e-3, steps, steps, scheme='midpoint')
    assert rel_err(got32, want32) < 2e-5 and rel_err(got64, want64) < 2e-5
    assert rel_err(got32, got64) > 0                      # the carry matters (and is not silently ignored)


# ---------------------------------------------------------------------------------
# edge cases and error behaviour
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize('n', (8, 30, 50, 200))
def test_odd_sizes_use_generic_path(n):
  """N not a multiple of 4 (and tiny N) exercise the generic conv path; N=200 is the
  reference's own test size (integrate_test.py:129-198)."""
  from ddd1d_b200 import model
  for kind, variant in (('burgers', 'plain'), ('ks', 'godunov')):
    hp = G.product_hparams(kind, variant, n)
    oeq = G.oracle_equation(kind, variant, n)
    w = O.glorot_weights(oeq, O.NetSpec(), seed=n, last_layer_scale=0.1, bias_scale=0.1)
    u = G.smooth_rows(3, n, seed=n)
    assert rel_err(cpu(model.predict_coefficients(u, hp, w)), O.predict_coefficients(u, oeq, O.NetSpec(), w)) < RHS_TOL
    # 1/dx^n stencils amplify float32 rounding with N: bound by the float32 reference graph's own error
    assert_f32_faithful(cpu(model.predict_time_derivative(u, hp, w)),
                        O.predict_time_derivative(u, oeq, O.NetSpec(), w),
                        O.predict_time_derivative(u, oeq, O.NetSpec(), w, dtype=np.float64),
                        what='%s %s N=%d' % (kind, variant, n))


def test_fast_and_generic_conv_paths_agree(monkeypatch):
  from ddd1d_b200 import model
  n = 64
  hp = G.product_hparams('kdv', 'conservative', n)
  oeq = G.oracle_equation('kdv', 'conservative', n)
  w = O.glorot_weights(oeq, O.NetSpec(), seed=5, last_layer_scale=0.1, bias_scale=0.1)
  u = G.smooth_rows(4, n, seed=1)
  fast = cpu(model.predict_time_derivative(u, hp, w))
  model._SOLVERS.clear()
  monkeypatch.setenv('DDD1D_FORCE_GENERIC_CONV', '1')
  monkeypatch.setenv('DDD1D_NO_BULK_COPY', '1')
  slow = cpu(model.predict_time_derivative(u, hp, w))
  model._SOLVERS.clear()
  assert rel_err(fast, slow) < 2e-6


def test_empty_single_and_oversubscribed_batches():
  import torch
  from ddd1d_b200 import integrate, equations
  eq = equations.KdVEquation(64)
  solver = integrate.BatchIntegrator.baseline([eq], 1)
  empty = solver.integrate(np.zeros((0, 64), np.float32), 0.0, 1e-5, 4, 2)
  assert tuple(empty.shape) == (2, 0, 64)
  one = solver.integrate(eq.initial_value()[None], 0.0, 1e-5, 4, 2)
  many = solver.integrate(np.tile(eq.initial_value()[None], (5000, 1)), 0.0, 1e-5, 4, 2)
  assert torch.equal(many[:, 0], one[:, 0]) and torch.equal(many[:, 4999], one[:, 0])
  # num_steps not a multiple of save_every: floor(num_steps / save_every) snapshots
  assert tuple(solver.integrate(eq.initial_value()[None], 0.0, 1e-5, 5, 2).shape) == (2, 1, 64)


def test_divergence_is_data_not_an_error():
  """Unstable step: rows go non-finite, first_bad_step says when, integrate_times NaN-pads
  (integrate.py:161-167)."""
  from ddd1d_b200 import integrate, equations
  eqs = [equations.KSEquation(256, random_seed=s) for s in range(3)]   # dx=0.25: 16 dt/dx^4 = 41 >> 2.5
  solver = integrate.BatchIntegrator.baseline(eqs, 1)
  u0 = solver.initial_values()
  snaps, bad = solver.integrate(u0, 0.0, 1e-2, 200, 50, return_first_bad=True)   # dt far too large
  bad = cpu(bad)
  assert (bad >= 0).all() and (bad < 200).all()
  assert not np.isfinite(cpu(snaps[-1])).any()
  ds = solver.integrate_times(u0, times=np.linspace(0, 2, 5), dt=1e-2)
  y = np.asarray(ds['y'].data)
  assert np.isfinite(y[:, 0]).all() and np.isnan(y[:, -1]).all()
  stable = integrate.BatchIntegrator.baseline(eqs, 1).integrate(u0, 0.0, 1e-5, 20, 20, return_first_bad=True)
  assert (cpu(stable[1]) == -1).all()


def test_argument_errors_mirror_the_reference():
  from ddd1d_b200 import model, integrate, equations, runtime
  hp = G.product_hparams('burgers', 'plain', 32)
  oeq = G.oracle_equation('burgers', 'plain', 32)
  w = O.glorot_weights(oeq, O.NetSpec(), seed=0)
  with pytest.raises(ValueError):                      # model.py:53-56
    model.predict_coefficients(np.zeros((2, 48), np.float32), hp, w)
  with pytest.raises(ValueError):                      # wrong weight shapes
    model.predict_coefficients(np.zeros((2, 32), np.float32), hp, w[:-1])
  with pytest.raises(ValueError):                      # integrate.py:320-321
    integrate.integrate_weno(equations.BurgersEquation(32))
  with pytest.raises(ValueError):                      # more rows than forcing samples
    integrate.BatchIntegrator.baseline([equations.BurgersEquation(32)], 1).integrate(
        np.zeros((2, 32), np.float32), 0.0, 1e-3, 1)
  with pytest.raises(NotImplementedError):             # coefficient grids beyond the 11-slot window
    runtime.learned_solver(equations.BurgersEquation(32),
                           G.product_hparams('burgers', 'plain', 32, coefficient_grid_min_size=13),
                           O.glorot_weights(oeq, O.NetSpec(coefficient_grid_min_size=13), seed=0))


def test_host_buffer_entry_points():
  """ddd1d_integrate_host / ddd1d_rhs_host: NumPy in, NumPy out (copies inside the library)."""
  from ddd1d_b200 import integrate, equations
  eqs = [equations.BurgersEquation(64, random_seed=s) for s in range(4)]
  solver = integrate.BatchIntegrator.baseline(eqs, 1).solver
  u0 = G.smooth_rows(4, 64, seed=3)
  snaps, bad = solver.integrate_host(u0, 0.0, 1e-3, 10, 5)
  dev = solver.integrate(u0, 0.0, 1e-3, 10, 5)
  np.testing.assert_array_equal(snaps, cpu(dev))
  assert (bad == -1).all()
  r = solver.rhs_host(0.3, u0.astype(np.float64))
  np.testing.assert_array_equal(r, cpu(solver.rhs(0.3, u0)).astype(np.float64))
  assert solver.launch_count() >= 3


# ---------------------------------------------------------------------------------
# the reference's integrate_test.py scenarios on the drop-in surface
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize('kind,conservative,numerical_flux', (
    ('burgers', False, False), ('burgers', True, False), ('burgers', True, True), ('kdv', True, False)))
def test_integrate_exact_baseline_and_model(kind, conservative, numerical_flux):
  """integrate_test.py:56-127 with explicit weights instead of an in-test training run:
  dims, zero-mean exact solution, identical (resampled) initial conditions, and y_baseline
  equal to a stand-alone integrate_baseline run."""
  import json
  from ddd1d_b200 import integrate, equations, training, duckarray, runtime
  fine_points, factor = 128, 4
  hp = training.create_hparams(kind, conservative=conservative, numerical_flux=numerical_flux,
                               resample_factor=factor, num_layers=1, filter_size=32,
                               equation_kwargs=json.dumps({'num_points': fine_points}))
  _, coarse = equations.from_hparams(hp, random_seed=0)
  shapes = runtime.expected_layer_shapes(coarse, hp)
  rs = np.random.RandomState(0)
  weights = [(1e-3 * rs.randn(*s).astype(np.float32), np.zeros(s[2], np.float32)) for s in shapes]
  times = np.linspace(0, 0.02 if kind == 'kdv' else 0.2, 3)
  ds = integrate.integrate_exact_baseline_and_model(weights, hparams=hp, random_seed=0, times=times)
  y_exact, y_base, y_model = (np.asarray(ds[k].data) for k in ('y_exact', 'y_baseline', 'y_model'))
  assert y_exact.shape == (3, fine_points) and y_base.shape == (3, fine_points // factor) == y_model.shape
  assert abs(y_exact.mean(axis=1)).max() < 1e-3
  resample = duckarray.resample_mean if conservative else duckarray.subsample
  np.testing.assert_allclose(resample(y_exact[0], factor), y_base[0])
  np.testing.assert_allclose(resample(y_exact[0], factor), y_model[0])
  assert np.isfinite(y_model).all()
  ds2 = integrate.integrate_baseline(type(coarse)(fine_points // factor, resample_factor=factor, random_seed=0),
                                     times=times)
  np.testing.assert_allclose(y_base, np.asarray(ds2['y'].data), atol=1e-5)


def test_spectral_exact_and_warmup(golden):
  """integrate_test.py:157-167 (exact == spectral for KdV) against the reference's own
  SpectralDifferentiator trajectory, and the warm-up + resample branch of integrate()."""
  from ddd1d_b200 import integrate, equations
  g = golden('trajectories')
  eq = equations.KdVEquation(64, random_seed=0)
  exact = integrate.integrate_exact(eq, times=g['spectral_kdv/times'])
  spectral = integrate.integrate_spectral(eq, times=g['spectral_kdv/times'])
  np.testing.assert_allclose(np.asarray(exact['y'].data), np.asarray(spectral['y'].data), atol=1e-10)
  np.testing.assert_allclose(np.asarray(exact['y'].data), g['spectral_kdv/y'], rtol=0, atol=1e-9)
  coarse = equations.ConservativeKdVEquation(16, resample_factor=4, random_seed=0)
  ds = integrate.integrate_baseline(coarse, times=np.linspace(0, 0.01, 3), warmup=0.01)
  y = np.asarray(ds['y'].data)
  assert y.shape == (3, 16) and np.isfinite(y).all()
  np.testing.assert_allclose(np.asarray(ds['time'].data if hasattr(ds['time'], 'data') else ds['time']),
                             0.01 + np.linspace(0, 0.01, 3))


def test_model_targets_against_reference_fixture(golden):
  """The other hparams.model_target values: the net predicts derivatives, dy/dt or a flux
  directly (model.py:551-640)."""
  from ddd1d_b200 import model, integrate
  g = golden('targets')
  for target in ('space_derivatives', 'time_derivative', 'flux'):
    for kind, variant in (('burgers', 'plain'), ('burgers', 'conservative'), ('ks', 'godunov')):
      key = '%s/%s/%s' % (target, kind, variant)
      hp = G.product_hparams(kind, variant, 32, model_target=target)
      w = weights_from(g, key)
      u = g[key + '/u']
      assert rel_err(cpu(model.predict_time_derivative(u, hp, w)), g[key + '/time_derivative']) < FLUX_TOL, key
      if target == 'space_derivatives':
        assert rel_err(cpu(model.predict_space_derivatives(u, hp, w)), g[key + '/space_derivatives']) < RHS_TOL
      else:
        with pytest.raises(NotImplementedError):
          model.predict_space_derivatives(u, hp, w)
      with pytest.raises(ValueError):
        model.predict_coefficients(u, hp, w)
      d = integrate.SavedModelDifferentiator(w, G.product_equation(kind, variant, 32, seed=5), hp)
      assert rel_err(d(0.61, u[0].astype(np.float64)), g[key + '/differentiator']) < FLUX_TOL, key


# ---------------------------------------------------------------------------------
# checkpoint directory ingestion (model.ckpt.* + hparams.pbtxt), SURVEY 8f rank 1
# ---------------------------------------------------------------------------------
def test_checkpoint_directory_drives_the_reference_surface(golden, tmp_path):
  """SavedModelDifferentiator(checkpoint_dir, ...) and the hparams=None wrappers read the
  TF-1 bundle + hparams.pbtxt and give exactly what explicit weights give (integrate.py:48-71,
  342-427); the fixture value is the reference's own graph output."""
  from ddd1d_b200 import checkpoint, integrate, training
  g = golden('learned')
  kind, variant, n = 'burgers', 'plain', 32
  key = 'default/%s/%s/%d' % (kind, variant, n)
  hp = G.product_hparams(kind, variant, n)
  w = weights_from(g, key)
  d = str(tmp_path / 'model_dir')
  checkpoint.save_conv_weights(d, w)
  training.save_hparams(d, hp)
  eq = G.product_equation(kind, variant, n, seed=7)
  u = g[key + '/u']
  from_dir = integrate.SavedModelDifferentiator(training.checkpoint_dir_to_path(d), eq, training.load_hparams(d))
  got = from_dir(float(g[key + '/t']), u[0].astype(np.float64))
  assert rel_err(got, g[key + '/differentiator']) < RHS_TOL
  explicit = integrate.SavedModelDifferentiator(w, eq, hp)
  np.testing.assert_array_equal(got, explicit(float(g[key + '/t']), u[0].astype(np.float64)))
  # hparams=None: read from the directory (integrate.py:406-407)
  times = np.linspace(0, 0.1, 3)
  y0 = u[0].astype(np.float64)
  a = integrate.integrate_model_from_warm_start(d, y0, times=times)
  b = integrate.integrate_model_from_warm_start(w, y0, hparams=hp, times=times)
  np.testing.assert_array_equal(np.asarray(a['y'].data), np.asarray(b['y'].data))


@pytest.mark.parametrize('kind,variant', [('burgers', 'plain'), ('burgers', 'conservative'),
                                          ('kdv', 'godunov'), ('ks', 'plain')])
def test_num_layers_zero_against_reference_fixture(golden, kind, variant):
  """hparams.num_layers = 0 (model.py:496-502): constant learned stencils; fixture = reference graph."""
  from ddd1d_b200 import model
  g = golden('layers0')
  key = '%s/%s' % (kind, variant)
  hp = G.product_hparams(kind, variant, 32, num_layers=0)
  w = [g[key + '/vector']]
  u = g[key + '/u']
  assert rel_err(cpu(model.predict_coefficients(u, hp, w)), g[key + '/coefficients']) < RHS_TOL
  assert rel_err(cpu(model.predict_time_derivative(u, hp, w)), g[key + '/time_derivative']) < RHS_TOL


def test_run_integrate_batch_matches_per_sample_scipy(golden, tmp_path):
  """evaluation.run_integrate_batch (all seeds in one device launch) vs the reference's per-seed
  execution model (scripts/run_evaluation.py:152-174: SavedModelDifferentiator + SciPy odeint)."""
  from ddd1d_b200 import checkpoint, evaluation, integrate, training
  from ddd1d_b200 import equations as equations_lib
  g = golden('learned')
  kind, variant, n = 'burgers', 'plain', 32
  key = 'default/%s/%s/%d' % (kind, variant, n)
  hp = G.product_hparams(kind, variant, n)
  w = weights_from(g, key)
  d = str(tmp_path / 'model_dir')
  checkpoint.save_conv_weights(d, w)
  training.save_hparams(d, hp)
  y0 = G.smooth_rows(3, n, seed=12).astype(np.float64)
  times = np.linspace(0, 0.5, 6)
  res = evaluation.run_integrate_batch(d, None, y0, times, warmup=1.0)
  assert res['y'].shape == (3, 6, n) and list(res['sample']) == [0, 1, 2]
  np.testing.assert_allclose(res['time'], 1.0 + times)
  for seed in range(3):
    _, coarse = equations_lib.from_hparams(hp, random_seed=seed)
    diff = integrate.SavedModelDifferentiator(training.checkpoint_dir_to_path(d), coarse, hp)
    want, nfev = integrate.odeint(y0[seed], diff, 1.0 + times)
    assert int(res['num_evals'][seed]) == nfev
    assert np.abs(res['y'][seed] - want).max() < 5e-5
  path = str(tmp_path / 'results.npz')
  evaluation.write_results(path, res)
  np.testing.assert_array_equal(evaluation.read_results(path)['y'], res['y'])


# ---------------------------------------------------------------------------------
# warp-per-row integrator of the modes without a net (csrc/ddd1d_warp.cuh)
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize('n', (32, 64, 128, 256))
@pytest.mark.parametrize('mode', ('fd', 'weno'))
def test_warp_row_kernel_against_oracle(n, mode):
  kinds = (('burgers', 1e-3), ('kdv', 2.5e-5), ('ks', 1e-5))
  for kind, dt in kinds:
    for variant in (('godunov',) if mode == 'weno' else VARIANTS):
      _fixed_step_case(kind, variant, n, 3, 24, dt, mode, tol=2e-4 if mode == 'weno' else TRAJ_TOL)


@pytest.mark.parametrize('scheme', ('rk3', 'midpoint', 'euler', 'rk4'))
def test_warp_row_kernel_equals_block_kernel(monkeypatch, scheme):
  """Same arithmetic, different thread mapping: the warp-per-row kernel and the CTA-per-row kernel agree to
  rounding on every equation variant, with forcing, snapshots, odd batch sizes and divergence reporting."""
  from ddd1d_b200 import integrate
  for kind, variant, mode, dt in (('burgers', 'plain', 'fd', 1e-3), ('burgers', 'conservative', 'fd', 1e-3),
                                  ('burgers', 'godunov', 'weno', 1e-3), ('ks', 'godunov', 'fd', 1e-5),
                                  ('kdv', 'conservative', 'fd', 2.5e-5)):
    n, batch = 64, 37
    eqs = [G.product_equation(kind, variant, n, seed=s) for s in range(batch)]
    solver = integrate.BatchIntegrator.weno(eqs) if mode == 'weno' else integrate.BatchIntegrator.baseline(eqs, 3)
    u0 = G.smooth_rows(batch, n, seed=4)
    if kind == 'burgers' and variant == 'plain':
      u0[5] *= 1e20                                     # this row blows up: divergence is data
    monkeypatch.delenv('DDD1D_NO_WARP_ROWS', raising=False)
    assert solver.solver.launch_shape(batch)['block'] == 256
    a, bad_a = solver.integrate(u0, 0.3, dt, 30, 10, scheme, return_first_bad=True)
    monkeypatch.setenv('DDD1D_NO_WARP_ROWS', '1')
    b, bad_b = solver.integrate(u0, 0.3, dt, 30, 10, scheme, return_first_bad=True)
    monkeypatch.delenv('DDD1D_NO_WARP_ROWS', raising=False)
    a, b = cpu(a), cpu(b)
    np.testing.assert_array_equal(cpu(bad_a), cpu(bad_b))
    ok = np.isfinite(b).all(axis=(0, 2))
    assert ok.sum() >= batch - 1
    assert rel_err(a[:, ok], b[:, ok]) < 2e-6, (kind, variant, mode, scheme)
    np.testing.assert_array_equal(np.isfinite(a), np.isfinite(b))


# ---------------------------------------------------------------------------------
# stand-alone helpers and stacked results (layers.py:95-100, model.py:162-275, 551-615, 664-696)
# ---------------------------------------------------------------------------------
def test_nn_conv1d_periodic_and_result_helpers(golden):
  import torch
  from ddd1d_b200 import layers, model
  g = golden('layers')
  x = np.arange(5.0, dtype=np.float32)[None, :, None]
  for name, filt in (('identity3', [0., 1., 0.]), ('shift2', [0., 1.]), ('avg2', [.5, .5]),
                     ('k4', [1., 2., 3., 4.]), ('k5', [1., 2., 3., 4., 5.])):
    f = np.array(filt, dtype=np.float32)[:, None, None]
    got = cpu(layers.nn_conv1d_periodic(x, f, center=True))[0, :, 0]
    np.testing.assert_allclose(got, g['conv/' + name], rtol=1e-6, atol=1e-6)
    # center=False: the window starts at the output point (layers.py:70-83)
    uncentred = cpu(layers.nn_conv1d_periodic(x, f, center=False))[0, :, 0]
    want = sum(filt[i] * np.roll(x[0, :, 0], -i) for i in range(len(filt)))
    np.testing.assert_allclose(uncentred, want, rtol=1e-6, atol=1e-6)

  # baseline_result = [space derivatives | time derivative | midpoint evolution] (model.py:245-275)
  gb = golden('baseline')
  eq = G.product_equation('burgers', 'plain', 32, seed=11)
  u = gb['burgers/plain/32/u']
  res = model.baseline_result(u, eq, num_time_steps=3, accuracy_order=1)
  sd, td, sol = model.result_unstack(res, eq)
  assert tuple(res.shape) == u.shape + (2 + 1 + 3,)
  assert rel_err(cpu(sd), gb['burgers/plain/32/acc1/space_derivatives']) < RHS_TOL
  assert rel_err(cpu(td), gb['burgers/plain/32/acc1/time_derivative']) < RHS_TOL
  oeq = G.oracle_equation('burgers', 'plain', 32, seed=11)
  def rhs(t, y):                                          # no forcing inside the training unroll (model.py:177-181)
    y32 = np.asarray(y, dtype=np.float32)
    return O.apply_space_derivatives(O.baseline_space_derivatives(y32, oeq, 1), y32, oeq).astype(np.float32)
  want = O.fixed_step_integrate(rhs, u, 0.0, oeq.time_step, 3, 1, scheme='midpoint')
  assert rel_err(cpu(sol), np.transpose(want, (1, 2, 0))) < TRAJ_TOL

  # the direct model targets through their own entry points (model.py:571-615) and predict_result
  gt = golden('targets')
  for target, fn in (('space_derivatives', model.predict_space_derivatives_directly),
                     ('time_derivative', model.predict_time_derivative_directly),
                     ('flux', model.predict_flux_directly)):
    key = '%s/burgers/plain' % target
    hp = G.product_hparams('burgers', 'plain', 32)              # model_target left at 'coefficients'
    w = weights_from(gt, key)
    want_key = key + ('/space_derivatives' if target == 'space_derivatives' else '/time_derivative')
    assert rel_err(cpu(fn(gt[key + '/u'], hp, w)), gt[want_key]) < RHS_TOL
  key = 'space_derivatives/burgers/plain'
  hp = G.product_hparams('burgers', 'plain', 32, model_target='space_derivatives')
  res = model.predict_result(gt[key + '/u'], hp, weights_from(gt, key))
  sd, td, sol = model.result_unstack(res, eq)
  assert sol is None
  assert rel_err(cpu(sd), gt[key + '/space_derivatives']) < RHS_TOL
  assert rel_err(cpu(td), gt[key + '/time_derivative']) < RHS_TOL
