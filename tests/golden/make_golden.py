"""Mint the golden fixtures in tests/golden/*.npz from the REFERENCE'S OWN CODE.

Run in the authoring container only (needs /root/reference):

    cd tests/golden && python make_golden.py

The reference's NumPy branches run as they are; its TensorFlow-graph code runs on
top of _tf_numpy_shim.py (a NumPy-eager stand-in for the primitive TF ops, see
that file's header for exactly which semantics are ours).  SciPy's solve_ivp is
driven through the reference's own integrate.odeint.  Nothing here is imported
by the gpu tests, smoke() or bench.py; the .npz files travel, this script's
inputs do not.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import _reference_import as R  # noqa: E402
import _tf_numpy_shim as S  # noqa: E402

(model, training, equations, layers, polynomials, weno, duckarray, integrate) = R.load(
    'model', 'training', 'equations', 'layers', 'polynomials', 'weno', 'duckarray',
    'integrate', numpy_tf=True)
import tensorflow as tf  # noqa: E402  (the shim)

VARIANTS = [('plain', False, False), ('conservative', True, False), ('godunov', True, True)]
KINDS = ['burgers', 'kdv', 'ks']


def equation_class(kind, variant):
  table = {'plain': equations.EQUATION_TYPES,
           'conservative': equations.CONSERVATIVE_EQUATION_TYPES,
           'godunov': equations.FLUX_EQUATION_TYPES}[variant]
  return table[kind]


def make_hparams(kind, variant, num_points, resample_factor=1, **overrides):
  cons = variant != 'plain'
  flux = variant == 'godunov'
  return training.create_hparams(
      kind, conservative=cons, numerical_flux=flux, resample_factor=resample_factor,
      equation_kwargs=json.dumps({'num_points': num_points * resample_factor}), **overrides)


def random_weights(shapes, seed, last_scale=0.1, bias_scale=0.1):
  rs = np.random.RandomState(seed)
  out = []
  for i, (k, cin, cout) in enumerate(shapes):
    limit = np.sqrt(6.0 / (k * cin + k * cout))
    w = rs.uniform(-limit, limit, size=(k, cin, cout))
    if i == len(shapes) - 1:
      w *= last_scale
    b = bias_scale * rs.uniform(-1, 1, size=(cout,))
    out.append((w.astype(np.float32), b.astype(np.float32)))
  return out


def conv_shapes(hp, equation):
  """Kernel shapes the reference's predict_coefficients will ask for (model.py:455-495)."""
  grid = polynomials.regular_grid(equation.GRID_OFFSET, 0, hp.coefficient_grid_min_size,
                                  equation.grid.solution_dx)
  if hp.model_target == 'space_derivatives':
    cout = len(equation.DERIVATIVE_ORDERS)
  elif hp.model_target in ('time_derivative', 'flux'):
    cout = 1
  elif hp.polynomial_accuracy_order:
    method = (polynomials.Method.FINITE_VOLUMES if equation.CONSERVATIVE
              else polynomials.Method.FINITE_DIFFERENCES)
    cout = sum(
        polynomials.PolynomialAccuracyLayer(grid, method, o, hp.polynomial_accuracy_order).input_size
        for o in equation.DERIVATIVE_ORDERS)
  else:
    cout = len(equation.DERIVATIVE_ORDERS) * grid.size
  shapes, cin = [], 1
  for _ in range(hp.num_layers - 1):
    shapes.append((hp.kernel_size, cin, hp.filter_size))
    cin = hp.filter_size
  shapes.append((hp.kernel_size, cin, cout))
  return shapes


def set_store(weights):
  S.STORE.pending = [(k.copy(), b.copy()) for k, b in weights]


def flat_weights(prefix, weights):
  out = {}
  for i, (k, b) in enumerate(weights):
    out['%s/kernel%d' % (prefix, i)] = k
    out['%s/bias%d' % (prefix, i)] = b
  return out


# ---------------------------------------------------------------------------
# eager twins of the graph-mode Differentiators (same op sequence as
# integrate.py:56-64 and :81-92, executed immediately instead of via sess.run)
# ---------------------------------------------------------------------------


class EagerModelDifferentiator(object):
  def __init__(self, equation, hparams, weights):
    self.equation, self.hparams, self.weights = equation, hparams, weights

  def __call__(self, t, y):
    inputs = tf.Tensor(np.asarray(y, dtype=np.float32))       # float32 placeholder feed
    set_store(self.weights)
    td = tf.squeeze(model.predict_time_derivative(inputs[tf.newaxis, :], self.hparams), axis=0)
    value = self.equation.finalize_time_derivative(tf.Tensor(np.float32(t)), td)
    return np.array(value.a)


class EagerPolynomialDifferentiator(object):
  def __init__(self, equation, accuracy_order=1):
    self.equation, self.accuracy_order = equation, accuracy_order

  def _derivs(self, y):
    inputs = tf.Tensor(np.asarray(y, dtype=np.float32))
    batched = inputs[tf.newaxis, :]
    return batched, model.baseline_space_derivatives(batched, self.equation, self.accuracy_order)

  def __call__(self, t, y):
    batched, sd = self._derivs(y)
    td = tf.squeeze(model.apply_space_derivatives(sd, batched, self.equation), axis=0)
    value = self.equation.finalize_time_derivative(tf.Tensor(np.float32(t)), td)
    return np.array(value.a)

  def calculate_space_derivatives(self, y):
    _, sd = self._derivs(y)
    return {k: np.array(tf.squeeze(sd[..., i], axis=0).a)
            for i, k in enumerate(self.equation.DERIVATIVE_NAMES)}


def reference_weno_differentiator(equation, non_weno_accuracy_order=3):
  d = object.__new__(integrate.WENODifferentiator)     # skip the graph-building __init__
  d.equation = equation
  d.poly_diff = EagerPolynomialDifferentiator(equation, non_weno_accuracy_order)
  return d


# ---------------------------------------------------------------------------


def golden_tables():
  out = {}
  for kind in KINDS:
    for variant, _, _ in VARIANTS:
      cls = equation_class(kind, variant)
      for n in (32, 64, 256):
        eq = cls(n)
        dx = eq.grid.solution_dx
        method = (polynomials.Method.FINITE_VOLUMES if eq.CONSERVATIVE
                  else polynomials.Method.FINITE_DIFFERENCES)
        grid = polynomials.regular_grid(eq.GRID_OFFSET, 0, 6, dx)
        key = '%s/%s/%d' % (kind, variant, n)
        out[key + '/grid'] = grid
        for d, order in enumerate(eq.DERIVATIVE_ORDERS):
          layer = polynomials.PolynomialAccuracyLayer(grid, method, order, 1)
          out['%s/bias%d' % (key, d)] = layer.bias
          out['%s/nullspace%d' % (key, d)] = layer.nullspace
          for acc in (1, 3):
            g = polynomials.regular_grid(eq.GRID_OFFSET, order, acc, dx)
            out['%s/fdgrid%d_acc%d' % (key, d, acc)] = g
            out['%s/fdcoef%d_acc%d' % (key, d, acc)] = polynomials.coefficients(g, method, order)
  np.savez_compressed(os.path.join(HERE, 'tables.npz'), **out)
  return len(out)


def golden_learned():
  out = {}
  rs = np.random.RandomState(1234)
  for kind in KINDS:
    for variant, _, _ in VARIANTS:
      for n in (32, 64):
        hp = make_hparams(kind, variant, n)
        eq = equation_class(kind, variant)(n, random_seed=7)
        weights = random_weights(conv_shapes(hp, eq), seed=len(out))
        u = (0.6 * rs.randn(3, n)).astype(np.float32)
        key = 'default/%s/%s/%d' % (kind, variant, n)
        out.update(flat_weights(key, weights))
        out[key + '/u'] = u
        set_store(weights)
        out[key + '/coefficients'] = model.predict_coefficients(tf.Tensor(u), hp).a
        set_store(weights)
        out[key + '/space_derivatives'] = model.predict_space_derivatives(tf.Tensor(u), hp).a
        set_store(weights)
        out[key + '/time_derivative'] = model.predict_time_derivative(tf.Tensor(u), hp).a
        # with finalize_time_derivative (forcing for Burgers), sample 0 only
        d = EagerModelDifferentiator(eq, hp, weights)
        out[key + '/t'] = np.float64(0.37)
        out[key + '/differentiator'] = d(0.37, u[0].astype(np.float64))
  # hparam variants on plain Burgers N=32 (and conservative for kernel_size 3)
  variants = {
      'num_layers1': dict(num_layers=1),
      'tanh_f16': dict(nonlinearity='tanh', filter_size=16),
      'elu_k3': dict(nonlinearity='elu', kernel_size=3),
      'softplus_l4': dict(nonlinearity='softplus', num_layers=4, filter_size=8),
      'relu6_scale': dict(nonlinearity='relu6', polynomial_accuracy_scale=0.5),
      'acc2': dict(polynomial_accuracy_order=2),
      'acc0': dict(polynomial_accuracy_order=0),
      'acc0_unbiased': dict(polynomial_accuracy_order=0, ensure_unbiased_coefficients=True),
  }
  for name, overrides in variants.items():
    for variant in ('plain', 'conservative'):
      if variant == 'conservative' and 'unbiased' in name:
        continue  # model.py:470-472 raises for 0th-order derivatives
      n = 32
      hp = make_hparams('burgers', variant, n, **overrides)
      eq = equation_class('burgers', variant)(n, random_seed=3)
      weights = random_weights(conv_shapes(hp, eq), seed=1000 + len(out))
      u = (0.6 * rs.randn(2, n)).astype(np.float32)
      key = '%s/burgers/%s/%d' % (name, variant, n)
      out[key + '/hparams'] = json.dumps(overrides)
      out.update(flat_weights(key, weights))
      out[key + '/u'] = u
      set_store(weights)
      out[key + '/coefficients'] = model.predict_coefficients(tf.Tensor(u), hp).a
      set_store(weights)
      out[key + '/time_derivative'] = model.predict_time_derivative(tf.Tensor(u), hp).a
  np.savez_compressed(os.path.join(HERE, 'learned.npz'), **out)
  return len(out)


def golden_targets():
  """The other hparams.model_target values (model.py:551-640), reference code on the shim."""
  out = {}
  rs = np.random.RandomState(4321)
  for target in ('space_derivatives', 'time_derivative', 'flux'):
    for kind, variant in (('burgers', 'plain'), ('burgers', 'conservative'), ('ks', 'godunov')):
      n = 32
      hp = make_hparams(kind, variant, n, model_target=target)
      eq = equation_class(kind, variant)(n, random_seed=5)
      weights = random_weights(conv_shapes(hp, eq), seed=len(out), last_scale=1.0)
      u = (0.6 * rs.randn(3, n)).astype(np.float32)
      key = '%s/%s/%s' % (target, kind, variant)
      out.update(flat_weights(key, weights))
      out[key + '/u'] = u
      set_store(weights)
      out[key + '/time_derivative'] = model.predict_time_derivative(tf.Tensor(u), hp).a
      if target == 'space_derivatives':
        set_store(weights)
        out[key + '/space_derivatives'] = model.predict_space_derivatives(tf.Tensor(u), hp).a
      d = EagerModelDifferentiator(eq, hp, weights)
      out[key + '/differentiator'] = d(0.61, u[0].astype(np.float64))
  np.savez_compressed(os.path.join(HERE, 'targets.npz'), **out)
  return len(out)


def golden_baseline():
  out = {}
  rs = np.random.RandomState(99)
  for kind in KINDS:
    for variant, _, _ in VARIANTS:
      n = 32
      eq = equation_class(kind, variant)(n, random_seed=11)
      u = (0.6 * rs.randn(3, n)).astype(np.float32)
      key = '%s/%s/%d' % (kind, variant, n)
      out[key + '/u'] = u
      for acc in (1, 3):
        sd = model.baseline_space_derivatives(tf.Tensor(u), eq, accuracy_order=acc)
        td = model.apply_space_derivatives(sd, tf.Tensor(u), eq)
        out['%s/acc%d/space_derivatives' % (key, acc)] = sd.a
        out['%s/acc%d/time_derivative' % (key, acc)] = td.a
        d = EagerPolynomialDifferentiator(eq, acc)
        out['%s/acc%d/differentiator' % (key, acc)] = d(1.25, u[0].astype(np.float64))
  # the "exact" dispatch of baseline_space_derivatives for Godunov Burgers (WENO in float32)
  eq = equations.GodunovBurgersEquation(32, random_seed=11)
  u = (0.6 * rs.randn(3, 32)).astype(np.float32)
  sd = model.baseline_space_derivatives(tf.Tensor(u), eq, accuracy_order=None)
  out['burgers/godunov/32/exact/u'] = u
  out['burgers/godunov/32/exact/space_derivatives'] = sd.a
  out['burgers/godunov/32/exact/time_derivative'] = model.apply_space_derivatives(sd, tf.Tensor(u), eq).a
  np.savez_compressed(os.path.join(HERE, 'baseline.npz'), **out)
  return len(out)


def golden_pointwise():
  """equations.py / weno.py / duckarray.py NumPy branches, float64."""
  out = {}
  rs = np.random.RandomState(5)
  n = 24
  y = rs.randn(2, n)
  out['y'] = y
  for kind in KINDS:
    for variant, _, _ in VARIANTS:
      eq = equation_class(kind, variant)(n, random_seed=2)
      derivs = {name: rs.randn(2, n) for name in eq.DERIVATIVE_NAMES}
      key = '%s/%s' % (kind, variant)
      for name, v in derivs.items():
        out['%s/deriv/%s' % (key, name)] = v
      out[key + '/equation_of_motion'] = eq.equation_of_motion(y, derivs)
      out[key + '/initial_value'] = eq.initial_value()
      out[key + '/time_step'] = eq.time_step
      out[key + '/standard_deviation'] = eq.standard_deviation
  um, up = rs.randn(50), rs.randn(50)
  out['godunov/u_minus'], out['godunov/u_plus'] = um, up
  out['godunov/flux'] = equations.godunov_convective_flux(um, up)
  out['staggered/y'] = y
  out['staggered/dy'] = equations.staggered_first_derivative(y, 0.3)
  # forcing: seeds x times x (resample factor, conservative)
  for seed in (0, 1, 17):
    for factor, cls in ((1, equations.BurgersEquation), (4, equations.BurgersEquation),
                        (4, equations.ConservativeBurgersEquation)):
      eq = cls(16, resample_factor=factor, random_seed=seed)
      key = 'forcing/%d/%d/%d' % (seed, factor, int(eq.CONSERVATIVE))
      out[key + '/a'] = eq.forcing.a
      out[key + '/omega'] = eq.forcing.omega
      out[key + '/k'] = eq.forcing.k
      out[key + '/phi'] = eq.forcing.phi
      for t in (0.0, 0.731, 12.5):
        out['%s/t%g/f64' % (key, t)] = eq.forcing(t)
        out['%s/t%g/f32' % (key, t)] = eq.forcing(tf.Tensor(np.float32(t))).a
  for seed in (0, 5):
    out['kdv_initial/%d' % seed] = equations.KdVEquation(32, random_seed=seed).initial_value()
    out['ks_initial/%d' % seed] = equations.KSEquation(32, random_seed=seed).initial_value()
    out['kdv_initial_cons_r4/%d' % seed] = equations.ConservativeKdVEquation(
        16, resample_factor=4, random_seed=seed).initial_value()
  # weno
  u = np.concatenate([rs.randn(2, 40), np.sign(rs.randn(2, 40))], axis=0)
  out['weno/u'] = u
  out['weno/left'] = weno.reconstruct_left(u)
  out['weno/right'] = weno.reconstruct_right(u)
  out['weno/omega'] = weno.calculate_omega(u)
  u32 = u.astype(np.float32)
  out['weno/left_f32'] = weno.reconstruct_left(tf.Tensor(u32)).a
  out['weno/right_f32'] = weno.reconstruct_right(tf.Tensor(u32)).a
  # resampling / spectral
  x = rs.randn(3, 48)
  out['resample/x'] = x
  out['resample/mean4'] = duckarray.resample_mean(x, 4)
  out['resample/sub4'] = duckarray.subsample(x, 4)
  out['spectral/d1'] = duckarray.spectral_derivative(x, 1, 7.0)
  out['spectral/d3'] = duckarray.spectral_derivative(x, 3, 7.0)
  out['spectral/filter'] = duckarray.smoothing_filter(x, order=4)
  np.savez_compressed(os.path.join(HERE, 'pointwise.npz'), **out)
  return len(out)


def golden_trajectories():
  out = {}
  # C1: Burgers, fixed polynomial FD coefficients (accuracy 1), N=64, T=2 ("200 RK steps")
  eq = equations.BurgersEquation(64)
  times = np.linspace(0, 2, 5)
  y, nfev = integrate.odeint(eq.initial_value(), EagerPolynomialDifferentiator(eq, 1), times)
  out['c1/times'], out['c1/y'], out['c1/nfev'] = times, y, nfev
  # other seeds (the batch axis of the GPU twin)
  for seed in (1, 2):
    eqs = equations.BurgersEquation(64, random_seed=seed)
    ys, nf = integrate.odeint(eqs.initial_value(), EagerPolynomialDifferentiator(eqs, 1), times)
    out['c1_seed%d/y' % seed], out['c1_seed%d/nfev' % seed] = ys, nf
  # conservative Burgers baseline, coarse grid with mean-resampled forcing
  eqc = equations.ConservativeBurgersEquation(32, resample_factor=4, random_seed=3)
  times_c = np.linspace(0, 1, 3)
  yc, nfc = integrate.odeint(eqc.initial_value(), EagerPolynomialDifferentiator(eqc, 1), times_c)
  out['cons_burgers/times'], out['cons_burgers/y'], out['cons_burgers/nfev'] = times_c, yc, nfc
  # learned model, Burgers N=32 (weights ~ small perturbation of the FD bias)
  n = 32
  hp = make_hparams('burgers', 'plain', n)
  eql = equations.BurgersEquation(n, random_seed=4)
  weights = random_weights(conv_shapes(hp, eql), seed=77, last_scale=0.01, bias_scale=0.0)
  tl = np.linspace(0, 0.5, 3)
  yl, nfl = integrate.odeint(eql.initial_value(), EagerModelDifferentiator(eql, hp, weights), tl)
  out.update(flat_weights('learned_burgers', weights))
  out['learned_burgers/times'], out['learned_burgers/y'], out['learned_burgers/nfev'] = tl, yl, nfl
  # learned KdV from its random initial condition (short: dt is stability limited)
  hpk = make_hparams('kdv', 'plain', n)
  eqk = equations.KdVEquation(n, random_seed=2)
  wk = random_weights(conv_shapes(hpk, eqk), seed=78, last_scale=0.01, bias_scale=0.0)
  tk = np.linspace(0, 0.05, 3)
  yk, nfk = integrate.odeint(eqk.initial_value(), EagerModelDifferentiator(eqk, hpk, wk), tk)
  out.update(flat_weights('learned_kdv', wk))
  out['learned_kdv/times'], out['learned_kdv/y'], out['learned_kdv/nfev'] = tk, yk, nfk
  # "exact" WENO Burgers (integrate.py:124-140 run by the reference's own __call__)
  eqw = equations.GodunovBurgersEquation(64, random_seed=1)
  tw = np.linspace(0, 1, 3)
  yw, nfw = integrate.odeint(eqw.initial_value(), reference_weno_differentiator(eqw), tw)
  out['weno_burgers/times'], out['weno_burgers/y'], out['weno_burgers/nfev'] = tw, yw, nfw
  dw = reference_weno_differentiator(eqw)
  uw = 0.5 * np.sin(eqw.grid.solution_x) + 0.1 * np.cos(3 * eqw.grid.solution_x)
  out['weno_burgers/rhs_u'] = uw
  out['weno_burgers/rhs'] = dw(0.4, uw)
  # spectral KdV (pure NumPy reference path, integrate.py:108-121)
  eqs = equations.KdVEquation(64, random_seed=0)
  ts = np.linspace(0, 0.02, 3)
  ys, nfs = integrate.odeint(eqs.initial_value(), integrate.SpectralDifferentiator(eqs), ts)
  out['spectral_kdv/times'], out['spectral_kdv/y'], out['spectral_kdv/nfev'] = ts, ys, nfs
  np.savez_compressed(os.path.join(HERE, 'trajectories.npz'), **out)
  return len(out)


def golden_layers():
  """layers.py run by the reference itself on the shim (alignment tables)."""
  out = {}
  for center in (True, False):
    for padding in (0, 1, 2, 3, 4, 5, 6, 7):
      x = tf.Tensor(np.arange(3.0, dtype=np.float32))[tf.newaxis, :, tf.newaxis]
      out['pad/%d/%d' % (int(center), padding)] = layers.pad_periodic(x, padding, center).a[0, :, 0]
  x = tf.Tensor(np.arange(5.0, dtype=np.float32))[tf.newaxis, :, tf.newaxis]
  for name, filt in (('identity3', [0., 1., 0.]), ('shift2', [0., 1.]), ('avg2', [.5, .5]),
                     ('k4', [1., 2., 3., 4.]), ('k5', [1., 2., 3., 4., 5.])):
    f = tf.Tensor(np.array(filt, dtype=np.float32))[:, tf.newaxis, tf.newaxis]
    out['conv/' + name] = layers.nn_conv1d_periodic(x, f, center=True).a[0, :, 0]
  np.savez_compressed(os.path.join(HERE, 'layers.npz'), **out)
  return len(out)


def golden_layers0():
  """hparams.num_layers = 0: one learned constant vector through the polynomial-accuracy layers
  (model.py:496-502), reference code on the shim."""
  out = {}
  rs = np.random.RandomState(2024)
  for kind, variant in (('burgers', 'plain'), ('burgers', 'conservative'), ('kdv', 'godunov'), ('ks', 'plain')):
    n = 32
    hp = make_hparams(kind, variant, n, num_layers=0)
    eq = equation_class(kind, variant)(n, random_seed=11)
    grid = polynomials.regular_grid(eq.GRID_OFFSET, 0, hp.coefficient_grid_min_size, eq.grid.solution_dx)
    method = (polynomials.Method.FINITE_VOLUMES if eq.CONSERVATIVE else polynomials.Method.FINITE_DIFFERENCES)
    size = sum(polynomials.PolynomialAccuracyLayer(grid, method, o, hp.polynomial_accuracy_order).input_size
               for o in eq.DERIVATIVE_ORDERS)
    vec = (0.1 * rs.randn(size)).astype(np.float32)
    u = (0.6 * rs.randn(2, n)).astype(np.float32)
    key = '%s/%s' % (kind, variant)
    out[key + '/vector'] = vec
    out[key + '/u'] = u
    S.STORE.named['coefficients'] = vec
    out[key + '/coefficients'] = model.predict_coefficients(tf.Tensor(u), hp).a
    out[key + '/time_derivative'] = model.predict_time_derivative(tf.Tensor(u), hp).a
  np.savez_compressed(os.path.join(HERE, 'layers0.npz'), **out)
  return len(out)


def golden_weno_parts():
  """The intermediate quantities of weno.py (smoothness indicators, omega, linear coefficients) and the
  index tables of layers.pad_periodic, from the reference's own NumPy / shim code."""
  out = {}
  rs = np.random.RandomState(77)
  u = rs.randn(3, 24)
  u[1] = np.where(np.arange(24) < 12, 1.0, -0.5)            # a discontinuity
  out['u'] = u
  out['indicators'] = weno.calculate_smoothness_indicators(u)
  out['omega'] = weno.calculate_omega(u)
  out['omega_reversed'] = weno.calculate_omega(u, weno.OPTIMAL_SMOOTH_WEIGHTS[::-1])
  out['left_coefficients'] = weno.left_coefficients(u)
  out['right_coefficients'] = weno.right_coefficients(u)
  out['left'] = weno.reconstruct_left(u)
  out['right'] = weno.reconstruct_right(u)
  x = np.arange(2 * 5 * 3, dtype=np.float32).reshape(2, 5, 3)
  out['pad/x'] = x
  for padding in (0, 1, 2, 3, 4, 6, 11, 13):
    for center in (False, True):
      out['pad/%d/%d' % (padding, int(center))] = layers.pad_periodic(tf.Tensor(x), padding, center=center).a
  np.savez_compressed(os.path.join(HERE, 'weno_parts.npz'), **out)
  return len(out)


def golden_grid9():
  """hparams.coefficient_grid_min_size = 9 (training_test.py:56 runs it on the default, conservative KS):
  9 centred points for the plain forms, 10 staggered points for the conservative / Godunov ones."""
  out = {}
  rs = np.random.RandomState(909)
  for kind, variant in (('ks', 'conservative'), ('ks', 'plain'), ('burgers', 'plain'), ('burgers', 'conservative'),
                        ('kdv', 'godunov')):
    for n in (32, 64):
      hp = make_hparams(kind, variant, n, coefficient_grid_min_size=9)
      eq = equation_class(kind, variant)(n, random_seed=9)
      weights = random_weights(conv_shapes(hp, eq), seed=500 + len(out))
      u = (0.6 * rs.randn(3, n)).astype(np.float32)
      key = '%s/%s/%d' % (kind, variant, n)
      out.update(flat_weights(key, weights))
      out[key + '/u'] = u
      set_store(weights)
      out[key + '/coefficients'] = model.predict_coefficients(tf.Tensor(u), hp).a
      set_store(weights)
      out[key + '/space_derivatives'] = model.predict_space_derivatives(tf.Tensor(u), hp).a
      set_store(weights)
      out[key + '/time_derivative'] = model.predict_time_derivative(tf.Tensor(u), hp).a
      d = EagerModelDifferentiator(eq, hp, weights)
      out[key + '/differentiator'] = d(0.23, u[0].astype(np.float64))
  np.savez_compressed(os.path.join(HERE, 'grid9.npz'), **out)
  return len(out)


if __name__ == '__main__':
  only = sys.argv[1:]
  for fn in (golden_tables, golden_learned, golden_targets, golden_baseline, golden_pointwise,
             golden_trajectories, golden_layers, golden_layers0, golden_weno_parts, golden_grid9):
    if only and fn.__name__ not in only:
      continue
    print(fn.__name__, fn())
  os.system('ls -la %s/*.npz' % HERE)
