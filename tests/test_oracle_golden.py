"""The oracle restatement vs fixtures minted from the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import pde_oracle as O
from tests.helpers import KINDS, VARIANTS, net_from_json, rel_err, weights_from

# float32 graphs: one rounding per op in the fixture vs NumPy float32 arithmetic here
F32_TOL = 2e-6


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('variant', VARIANTS)
def test_tables(golden, kind, variant):
  g = golden('tables')
  for n in (32, 64, 256):
    eq = O.EquationSpec(kind, variant, num_points=n)
    net = O.NetSpec()
    key = '%s/%s/%d' % (kind, variant, n)
    np.testing.assert_allclose(O.coefficient_grid(eq, net), g[key + '/grid'], rtol=0, atol=1e-15)
    method = O.FINITE_VOLUMES if eq.conservative else O.FINITE_DIFFERENCES
    for d, (layer, order) in enumerate(zip(O.accuracy_layers(eq, net), eq.orders)):
      np.testing.assert_allclose(layer.bias, g['%s/bias%d' % (key, d)], rtol=1e-12, atol=1e-12)
      # same LAPACK, same matrix => same basis (sign included)
      np.testing.assert_allclose(layer.nullspace, g['%s/nullspace%d' % (key, d)], rtol=1e-10, atol=1e-9)
      for acc in (1, 3):
        grid = O.regular_grid(eq.grid_offset, order, acc, eq.dx)
        np.testing.assert_allclose(grid, g['%s/fdgrid%d_acc%d' % (key, d, acc)], rtol=0, atol=1e-15)
        np.testing.assert_allclose(O.coefficients(grid, method, order),
                                   g['%s/fdcoef%d_acc%d' % (key, d, acc)], rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('variant', VARIANTS)
@pytest.mark.parametrize('n', (32, 64))
def test_learned_default_net(golden, kind, variant, n):
  g = golden('learned')
  key = 'default/%s/%s/%d' % (kind, variant, n)
  eq = O.EquationSpec(kind, variant, num_points=n, random_seed=7)
  net = O.NetSpec()
  w = weights_from(g, key)
  u = g[key + '/u']
  coefs = O.predict_coefficients(u, eq, net, w)
  assert coefs.dtype == np.float32
  assert rel_err(coefs, g[key + '/coefficients']) < F32_TOL
  derivs = O.apply_coefficients(coefs, u)
  assert rel_err(derivs, g[key + '/space_derivatives']) < 5 * F32_TOL
  assert rel_err(O.predict_time_derivative(u, eq, net, w), g[key + '/time_derivative']) < 1e-5
  d = O.ModelDifferentiator(eq, net, w)
  assert rel_err(d(float(g[key + '/t']), u[0].astype(np.float64)), g[key + '/differentiator']) < 1e-5


def test_learned_hparam_variants(golden):
  g = golden('learned')
  names = sorted({k.split('/')[0] for k in g.files} - {'default'})
  assert len(names) == 8
  for name in names:
    for variant in ('plain', 'conservative'):
      key = '%s/burgers/%s/32' % (name, variant)
      if key + '/u' not in g.files:
        continue
      net = net_from_json(g[key + '/hparams'])
      eq = O.EquationSpec('burgers', variant, num_points=32, random_seed=3)
      w = weights_from(g, key)
      assert [k.shape for k, _ in w] == O.layer_shapes(eq, net)
      u = g[key + '/u']
      assert rel_err(O.predict_coefficients(u, eq, net, w), g[key + '/coefficients']) < F32_TOL, key
      assert rel_err(O.predict_time_derivative(u, eq, net, w), g[key + '/time_derivative']) < 1e-5, key


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('variant', VARIANTS)
def test_baseline(golden, kind, variant):
  g = golden('baseline')
  key = '%s/%s/32' % (kind, variant)
  eq = O.EquationSpec(kind, variant, num_points=32, random_seed=11)
  u = g[key + '/u']
  for acc in (1, 3):
    sd = O.baseline_space_derivatives(u, eq, acc)
    assert rel_err(sd, g['%s/acc%d/space_derivatives' % (key, acc)]) < F32_TOL
    td = O.apply_space_derivatives(sd, u, eq)
    assert rel_err(td, g['%s/acc%d/time_derivative' % (key, acc)]) < 1e-5
    d = O.PolynomialDifferentiator(eq, acc)
    assert rel_err(d(1.25, u[0].astype(np.float64)), g['%s/acc%d/differentiator' % (key, acc)]) < 1e-5


def test_baseline_exact_weno_float32(golden):
  g = golden('baseline')
  eq = O.EquationSpec('burgers', 'godunov', num_points=32, random_seed=11)
  u = g['burgers/godunov/32/exact/u']
  sd = O.baseline_space_derivatives(u, eq, None)
  assert rel_err(sd, g['burgers/godunov/32/exact/space_derivatives']) < 1e-5


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('variant', VARIANTS)
def test_equation_of_motion(golden, kind, variant):
  g = golden('pointwise')
  eq = O.EquationSpec(kind, variant, num_points=24, random_seed=2)
  key = '%s/%s' % (kind, variant)
  derivs = {name: g['%s/deriv/%s' % (key, name)] for name in eq.names}
  np.testing.assert_allclose(eq.equation_of_motion(g['y'], derivs), g[key + '/equation_of_motion'],
                             rtol=1e-13, atol=1e-13)
  np.testing.assert_allclose(eq.initial_value(), g[key + '/initial_value'], rtol=1e-13, atol=1e-14)
  assert eq.time_step == float(g[key + '/time_step'])
  assert eq.standard_deviation == float(g[key + '/standard_deviation'])


def test_pointwise_misc(golden):
  g = golden('pointwise')
  np.testing.assert_allclose(O.godunov_convective_flux(g['godunov/u_minus'], g['godunov/u_plus']),
                             g['godunov/flux'], rtol=0, atol=0)
  np.testing.assert_allclose(O.staggered_first_derivative(g['staggered/y'], 0.3), g['staggered/dy'],
                             rtol=1e-14)
  for seed in (0, 1, 17):
    for factor, cons in ((1, 0), (4, 0), (4, 1)):
      eq = O.EquationSpec('burgers', 'conservative' if cons else 'plain', num_points=16,
                          resample_factor=factor, random_seed=seed)
      key = 'forcing/%d/%d/%d' % (seed, factor, cons)
      for name in ('a', 'omega', 'k', 'phi'):
        np.testing.assert_array_equal(getattr(eq.forcing, name), g['%s/%s' % (key, name)])
      for t in (0.0, 0.731, 12.5):
        np.testing.assert_allclose(eq.forcing(t), g['%s/t%g/f64' % (key, t)], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(eq.forcing(t, dtype=np.float32), g['%s/t%g/f32' % (key, t)],
                                   rtol=0, atol=3e-6)
  for seed in (0, 5):
    np.testing.assert_allclose(O.EquationSpec('kdv', num_points=32, random_seed=seed).initial_value(),
                               g['kdv_initial/%d' % seed], rtol=1e-13)
    np.testing.assert_allclose(O.EquationSpec('ks', num_points=32, random_seed=seed).initial_value(),
                               g['ks_initial/%d' % seed], rtol=1e-13)
    np.testing.assert_allclose(
        O.EquationSpec('kdv', 'conservative', num_points=16, resample_factor=4,
                       random_seed=seed).initial_value(),
        g['kdv_initial_cons_r4/%d' % seed], rtol=1e-13, atol=1e-15)


def test_weno_and_resample(golden):
  g = golden('pointwise')
  u = g['weno/u']
  np.testing.assert_allclose(O.weno_reconstruct_left(u), g['weno/left'], rtol=1e-13, atol=1e-14)
  np.testing.assert_allclose(O.weno_reconstruct_right(u), g['weno/right'], rtol=1e-13, atol=1e-14)
  np.testing.assert_allclose(O.weno_omega(u), g['weno/omega'], rtol=1e-13)
  u32 = u.astype(np.float32)
  np.testing.assert_allclose(O.weno_reconstruct_left(u32), g['weno/left_f32'], rtol=0, atol=2e-5)
  np.testing.assert_allclose(O.weno_reconstruct_right(u32), g['weno/right_f32'], rtol=0, atol=2e-5)
  x = g['resample/x']
  np.testing.assert_allclose(O.resample_mean(x, 4), g['resample/mean4'], rtol=1e-14)
  np.testing.assert_array_equal(O.subsample(x, 4), g['resample/sub4'])
  np.testing.assert_allclose(O.spectral_derivative(x, 1, 7.0), g['spectral/d1'], rtol=1e-12, atol=1e-12)
  np.testing.assert_allclose(O.spectral_derivative(x, 3, 7.0), g['spectral/d3'], rtol=1e-12, atol=1e-10)
  np.testing.assert_allclose(O.smoothing_filter(x, order=4), g['spectral/filter'], rtol=1e-12, atol=1e-13)


def test_layers_alignment(golden):
  g = golden('layers')
  for center in (True, False):
    for padding in range(8):
      x = np.arange(3.0, dtype=np.float32)[None, :, None]
      np.testing.assert_array_equal(O.pad_periodic(x, padding, center)[0, :, 0],
                                    g['pad/%d/%d' % (int(center), padding)])
  x = np.arange(5.0, dtype=np.float32)[None, :, None]
  for name, filt in (('identity3', [0., 1., 0.]), ('shift2', [0., 1.]), ('avg2', [.5, .5]),
                     ('k4', [1., 2., 3., 4.]), ('k5', [1., 2., 3., 4., 5.])):
    f = np.array(filt, dtype=np.float32)[:, None, None]
    np.testing.assert_allclose(O.nn_conv1d_periodic(x, f, center=True)[0, :, 0], g['conv/' + name])


def test_trajectories(golden):
  """SciPy RK23 driven exactly like integrate.odeint (integrate.py:143-169)."""
  g = golden('trajectories')
  # C1 (BASELINE config 1): Burgers FD accuracy 1, N=64, T=2
  for tag, seed in (('c1', 0), ('c1_seed1', 1), ('c1_seed2', 2)):
    eq = O.EquationSpec('burgers', num_points=64, random_seed=seed)
    y, nfev = O.odeint(eq.initial_value(), O.PolynomialDifferentiator(eq, 1), g['c1/times'])
    assert nfev == int(g[tag + '/nfev'])
    np.testing.assert_allclose(y, g[tag + '/y'], rtol=0, atol=2e-5)
  eq = O.EquationSpec('burgers', 'conservative', num_points=32, resample_factor=4, random_seed=3)
  y, nfev = O.odeint(eq.initial_value(), O.PolynomialDifferentiator(eq, 1), g['cons_burgers/times'])
  assert nfev == int(g['cons_burgers/nfev'])
  np.testing.assert_allclose(y, g['cons_burgers/y'], rtol=0, atol=2e-5)
  eq = O.EquationSpec('burgers', num_points=32, random_seed=4)
  w = weights_from(g, 'learned_burgers')
  y, nfev = O.odeint(eq.initial_value(), O.ModelDifferentiator(eq, O.NetSpec(), w), g['learned_burgers/times'])
  assert nfev == int(g['learned_burgers/nfev'])
  np.testing.assert_allclose(y, g['learned_burgers/y'], rtol=0, atol=2e-5)
  eq = O.EquationSpec('kdv', num_points=32, random_seed=2)
  w = weights_from(g, 'learned_kdv')
  y, nfev = O.odeint(eq.initial_value(), O.ModelDifferentiator(eq, O.NetSpec(), w), g['learned_kdv/times'])
  assert nfev == int(g['learned_kdv/nfev'])
  np.testing.assert_allclose(y, g['learned_kdv/y'], rtol=0, atol=5e-5)
  eq = O.EquationSpec('burgers', 'godunov', num_points=64, random_seed=1)
  d = O.WENODifferentiator(eq)
  np.testing.assert_allclose(d(0.4, g['weno_burgers/rhs_u']), g['weno_burgers/rhs'], rtol=0, atol=1e-6)
  y, nfev = O.odeint(eq.initial_value(), d, g['weno_burgers/times'])
  assert nfev == int(g['weno_burgers/nfev'])
  np.testing.assert_allclose(y, g['weno_burgers/y'], rtol=0, atol=1e-6)
  eq = O.EquationSpec('kdv', num_points=64, random_seed=0)
  y, nfev = O.odeint(eq.initial_value(), O.SpectralDifferentiator(eq), g['spectral_kdv/times'])
  assert nfev == int(g['spectral_kdv/nfev'])
  np.testing.assert_allclose(y, g['spectral_kdv/y'], rtol=1e-12, atol=1e-12)


def test_model_targets(golden):
  """hparams.model_target in {space_derivatives, time_derivative, flux} (model.py:551-640)."""
  g = golden('targets')
  for target in ('space_derivatives', 'time_derivative', 'flux'):
    for kind, variant in (('burgers', 'plain'), ('burgers', 'conservative'), ('ks', 'godunov')):
      key = '%s/%s/%s' % (target, kind, variant)
      eq = O.EquationSpec(kind, variant, num_points=32, random_seed=5)
      net = O.NetSpec(model_target=target)
      w = weights_from(g, key)
      assert [k.shape for k, _ in w] == O.layer_shapes(eq, net)
      u = g[key + '/u']
      assert rel_err(O.predict_time_derivative(u, eq, net, w), g[key + '/time_derivative']) < 1e-5, key
      if target == 'space_derivatives':
        assert rel_err(O.multilayer_conv1d(u, eq, net, w), g[key + '/space_derivatives']) < F32_TOL
      d = O.ModelDifferentiator(eq, net, w)
      assert rel_err(d(0.61, u[0].astype(np.float64)), g[key + '/differentiator']) < 1e-5, key


@pytest.mark.parametrize('kind,variant', [('burgers', 'plain'), ('burgers', 'conservative'),
                                          ('kdv', 'godunov'), ('ks', 'plain')])
def test_num_layers_zero(golden, kind, variant):
  """hparams.num_layers = 0 (model.py:496-502): a learned constant vector instead of a net."""
  g = golden('layers0')
  key = '%s/%s' % (kind, variant)
  eq = O.EquationSpec(kind, variant, num_points=32, random_seed=11)
  net = O.NetSpec(num_layers=0)
  w = [g[key + '/vector']]
  u = g[key + '/u']
  assert rel_err(O.predict_coefficients(u, eq, net, w), g[key + '/coefficients']) < F32_TOL
  assert rel_err(O.predict_time_derivative(u, eq, net, w), g[key + '/time_derivative']) < F32_TOL


GRID9_CASES = (('ks', 'conservative'), ('ks', 'plain'), ('burgers', 'plain'), ('burgers', 'conservative'),
               ('kdv', 'godunov'))


@pytest.mark.parametrize('kind,variant', GRID9_CASES)
@pytest.mark.parametrize('n', (32, 64))
def test_coefficient_grid_min_size_9(golden, kind, variant, n):
  """hparams.coefficient_grid_min_size = 9 (model.py:445-448; training_test.py:56): 9 centred / 10 staggered points."""
  g = golden('grid9')
  key = '%s/%s/%d' % (kind, variant, n)
  eq = O.EquationSpec(kind, variant, num_points=n, random_seed=9)
  net = O.NetSpec(coefficient_grid_min_size=9)
  assert O.coefficient_grid(eq, net).size == (9 if variant == 'plain' else 10)
  w = weights_from(g, key)
  u = g[key + '/u']
  coefs = O.predict_coefficients(u, eq, net, w)
  assert coefs.shape == g[key + '/coefficients'].shape
  assert rel_err(coefs, g[key + '/coefficients']) < F32_TOL
  assert rel_err(O.apply_coefficients(coefs, u), g[key + '/space_derivatives']) < 5 * F32_TOL
  assert rel_err(O.predict_time_derivative(u, eq, net, w), g[key + '/time_derivative']) < 1e-5
  d = O.ModelDifferentiator(eq, net, w)
  assert rel_err(d(0.23, u[0].astype(np.float64)), g[key + '/differentiator']) < 1e-5
