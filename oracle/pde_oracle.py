"""CPU oracle for the pde_superresolution hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, in plain NumPy (+ SciPy's solve_ivp as the reference itself
uses), the algorithm of google/data-driven-discretization-1d's time-integration
path.  It is the *checker* for the CUDA implementation: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  The product package never does.

Parity pinning (see DESIGN.md, section "Oracle"):
  * host-side tables (regular_grid / constraints / coefficients / null space),
    pointwise equations, Godunov flux, forcing, WENO5 and resampling are pinned
    against the reference's OWN NumPy code, executed in the authoring container
    under a stand-in ``tensorflow`` module (tests/golden/make_golden.py writes the
    fixtures in tests/golden/*.npz), and against the reference's known-answer
    tests (layers_test.py, polynomials_test.py, weno_test.py, equations_test.py,
    duckarray_test.py) which tests/test_oracle_kat.py re-expresses without TF;
  * the TF-graph-only pieces (model.predict_coefficients / predict_time_derivative /
    baseline_space_derivatives, layers.pad_periodic, polynomials.reconstruct) cannot run on
    TensorFlow here (tensorflow<2 is not installable offline); they are pinned against the
    reference's own graph-construction code executed on a NumPy-eager stand-in for the primitive
    TF ops (tests/golden/_tf_numpy_shim.py: VALID conv1d, einsum, extract_image_patches, ... with
    TF's documented semantics).  What stays unpinned is only the float32 summation order of TF's
    real conv kernels; the reference itself holds no golden vector for the learned path
    (integrate_test.py:50-54 trains on noise inside the test and asserts only shapes).
    torch.nn.functional.conv1d on CPU is an independent second opinion in tests/test_oracle_kat.py.

All file:line citations are into /root/reference/pde_superresolution/.
"""
from __future__ import annotations

import math

import numpy as np

# ----------------------------------------------------------------------------
# layers.py
# ----------------------------------------------------------------------------


def pad_periodic(inputs, padding, center=False):
  """layers.py:39-83.  inputs [batch, length, features] -> [batch, length+padding, features].

  center=True puts ceil(padding/2) wrapped points on the left (python's
  ``-padding//2`` slice, layers.py:77) and floor(padding/2) on the right (:79);
  center=False appends ``padding`` points on the right (:81).  Padding larger
  than the period tiles the input (:70-74).
  """
  inputs = np.asarray(inputs)
  if inputs.ndim != 3:
    raise ValueError('inputs must be 3D for periodic padding')
  if padding == 0:
    return inputs
  n = inputs.shape[1]
  if center:
    left, right = -(-padding // 2), padding // 2
  else:
    left, right = 0, padding
  index = np.arange(-left, n + right) % n
  return inputs[:, index, :]


def nn_conv1d_periodic(inputs, filters, center=False):
  """layers.py:95-100: VALID cross-correlation after periodic padding.

  inputs [b, x, cin]; filters [k, cin, cout] (TF layout).
  """
  inputs = np.asarray(inputs)
  filters = np.asarray(filters)
  k = filters.shape[0]
  padded = pad_periodic(inputs, k - 1, center=center)
  n = inputs.shape[1]
  out = np.zeros(inputs.shape[:2] + (filters.shape[2],), dtype=np.result_type(inputs, filters))
  for tap in range(k):
    out += padded[:, tap:tap + n, :] @ filters[tap]
  return out


ACTIVATIONS = {
    # model.py:411-417
    'relu': lambda x: np.maximum(x, 0),
    'relu6': lambda x: np.minimum(np.maximum(x, 0), 6),
    'tanh': np.tanh,
    'softplus': lambda x: np.logaddexp(x, 0).astype(x.dtype),
    'elu': lambda x: np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(x.dtype),
    None: lambda x: x,
}


def conv1d_periodic_layer(inputs, kernel, bias, activation=None, center=True):
  """layers.py:103-137 (tf.layers.conv1d, padding='valid', stride 1, dilation 1)."""
  out = nn_conv1d_periodic(inputs, kernel, center=center) + bias
  return ACTIVATIONS[activation](out)


# ----------------------------------------------------------------------------
# polynomials.py (host-side table construction)
# ----------------------------------------------------------------------------

CENTERED, STAGGERED = 'centered', 'staggered'          # polynomials.py:31-34
FINITE_DIFFERENCES, FINITE_VOLUMES = 'fd', 'fv'         # polynomials.py:37-40


def regular_grid(grid_offset, derivative_order, accuracy_order=1, dx=1.0):
  """polynomials.py:43-71."""
  min_size = derivative_order + accuracy_order
  if grid_offset == CENTERED:
    m = min_size // 2
    return np.arange(-m, m + 1) * dx
  if grid_offset == STAGGERED:
    m = (min_size + 1) // 2
    return (0.5 + np.arange(-m, m)) * dx
  raise ValueError('unexpected grid_offset: {}'.format(grid_offset))


def constraints(grid, method, derivative_order, accuracy_order=None):
  """polynomials.py:74-149: rows of A are Taylor (FD) or cell-average (FV) moments."""
  grid = np.asarray(grid, dtype=float)
  if accuracy_order is None:
    accuracy_order = grid.size - derivative_order
  if accuracy_order < 1:
    raise ValueError('cannot compute constraints with non-positive accuracy_order')
  deltas = np.unique(np.diff(grid))
  if (abs(deltas - deltas[0]) > 1e-8).any():
    raise ValueError('not a regular grid: {}'.format(deltas))
  delta = deltas[0]
  final = None
  zero_rows = set()
  for m in range(accuracy_order + derivative_order):
    if method == FINITE_DIFFERENCES:
      row = grid ** m                                               # :123
    elif method == FINITE_VOLUMES:
      row = (1 / delta * ((grid + delta / 2) ** (m + 1)             # :124-128
                          - (grid - delta / 2) ** (m + 1)) / (m + 1))
    else:
      raise ValueError('unexpected method: {}'.format(method))
    if m == derivative_order:
      final = row
    else:
      zero_rows.add(tuple(row))                                     # dedup :134
  if len(zero_rows) + 1 > grid.size:
    raise ValueError('no valid stencil exists')
  a = np.array(sorted(zero_rows) + [final])                         # :144
  b = np.zeros(a.shape[0])
  b[-1] = math.factorial(derivative_order)                          # :147
  return a, b


def coefficients(grid, method, derivative_order):
  """polynomials.py:152-167."""
  a, b = constraints(grid, method, derivative_order)
  return np.linalg.solve(a, b)


def zero_padded_coefficients(grid, method, derivative_order, padding):
  """polynomials.py:170-195."""
  left, right = padding
  trimmed = np.asarray(grid)[left:(-right or None)]
  return np.pad(coefficients(trimmed, method, derivative_order), padding, mode='constant')


class PolynomialAccuracyLayer(object):
  """polynomials.py:198-277: coef = bias + z @ nullspace, A @ coef == b for every z."""

  def __init__(self, grid, method, derivative_order, accuracy_order=2,
               bias=None, bias_zero_padding=(0, 0), out_scale=1.0):
    grid = np.asarray(grid, dtype=float)
    a, b = constraints(grid, method, derivative_order, accuracy_order)
    if bias is None:
      bias = zero_padded_coefficients(grid, method, derivative_order, bias_zero_padding)
    if np.linalg.norm(a @ bias - b) > 1e-8:                          # :241-243
      raise ValueError('invalid bias, not in nullspace')
    _, _, v = np.linalg.svd(a)                                       # :246
    input_size = a.shape[1] - a.shape[0]
    if not input_size:
      raise ValueError('there is only one valid solution accurate to this order')
    dx = grid[1] - grid[0]
    self.input_size = input_size
    self.grid_size = grid.size
    self.nullspace = v[-input_size:] * (out_scale / dx ** derivative_order)   # :254-259
    self.bias = bias

  def apply(self, inputs):
    """polynomials.py:266-277 (float32 tables, einsum 'bxi,ij->bxj')."""
    dtype = inputs.dtype
    return self.bias.astype(dtype) + inputs @ self.nullspace.astype(dtype)


def reconstruct(inputs, grid, method, derivative_order):
  """polynomials.py:280-303: constant stencil, coefficients cast to the input dtype."""
  inputs = np.asarray(inputs)
  filt = coefficients(grid, method, derivative_order).astype(inputs.dtype)
  out = nn_conv1d_periodic(inputs[..., None], filt[:, None, None], center=True)
  return out[..., 0]


# ----------------------------------------------------------------------------
# duckarray.py
# ----------------------------------------------------------------------------


def resample_mean(x, factor, axis=-1):
  """duckarray.py:139-163."""
  x = np.asarray(x)
  axis = axis % x.ndim
  if x.shape[axis] % factor:
    raise ValueError('resample factor {} must divide size {}'.format(factor, x.shape[axis]))
  shape = x.shape[:axis] + (x.shape[axis] // factor, factor) + x.shape[axis + 1:]
  return x.reshape(shape).mean(axis=axis + 1)


def subsample(x, factor, axis=-1):
  """duckarray.py:166-189."""
  x = np.asarray(x)
  axis = axis % x.ndim
  if x.shape[axis] % factor:
    raise ValueError('resample factor {} must divide size {}'.format(factor, x.shape[axis]))
  index = [slice(None)] * x.ndim
  index[axis] = slice(None, None, factor)
  return x[tuple(index)]


def spectral_derivative(x, order=1, period=2 * np.pi):
  """duckarray.py:105-112."""
  n = x.shape[-1]
  if n % 2:
    raise ValueError('spectral derivative only works for even length data')
  c = 2j * np.pi / period
  k = np.fft.rfftfreq(n, d=1 / n)
  return np.fft.irfft((c * k) ** order * np.fft.rfft(x))


def smoothing_filter(x, alpha=-np.log(1e-15), order=2):
  """duckarray.py:115-128."""
  n = x.shape[-1]
  if n % 2:
    raise ValueError('smoothing filter only works for even length data')
  count = n // 2
  eta = np.arange(count + 1) / count
  sigma = np.exp(-alpha * eta ** (2 * order))
  return np.fft.irfft(sigma * np.fft.rfft(x))


# ----------------------------------------------------------------------------
# weno.py
# ----------------------------------------------------------------------------


def _roll(u, shift):
  return np.roll(u, shift, axis=-1)


def weno_smoothness(u):
  """weno.py:43-57."""
  um2, um1, up1, up2 = _roll(u, 2), _roll(u, 1), _roll(u, -1), _roll(u, -2)
  return np.stack([
      1 / 4 * (um2 - 4 * um1 + 3 * u) ** 2 + 13 / 12 * (um2 - 2 * um1 + u) ** 2,
      1 / 4 * (um1 - up1) ** 2 + 13 / 12 * (um1 - 2 * u + up1) ** 2,
      1 / 4 * (3 * u - 4 * up1 + up2) ** 2 + 13 / 12 * (u - 2 * up1 + up2) ** 2,
  ], axis=-2)


def weno_omega(u, linear_weights=(0.1, 0.6, 0.3), epsilon=1e-6, p=2):
  """weno.py:60-73."""
  beta = weno_smoothness(u)
  alpha = np.array(linear_weights)[:, None] / (epsilon + beta) ** p
  return alpha / alpha.sum(axis=-2, keepdims=True)


def weno_reconstruct_left(u):
  """weno.py:76-97: u at i+1/2 from the left-biased 5-point stencil."""
  w = weno_omega(u)
  w0, w1, w2 = w[..., 0, :], w[..., 1, :], w[..., 2, :]
  c = np.stack([w0 / 3, -(7 * w0 + w1) / 6, (11 * w0 + 5 * w1 + 2 * w2) / 6,
                (2 * w1 + 5 * w2) / 6, -w2 / 6], axis=-1)
  u_all = np.stack([_roll(u, i) for i in (2, 1, 0, -1, -2)], axis=-1)
  return (c * u_all).sum(axis=-1)


def weno_reconstruct_right(u):
  """weno.py:100-123 (reversed linear weights, omega rolled by -1)."""
  w = _roll(weno_omega(u, (0.3, 0.6, 0.1)), -1)
  w2, w1, w0 = w[..., 0, :], w[..., 1, :], w[..., 2, :]
  c = np.stack([-w2 / 6, (5 * w2 + 2 * w1) / 6, (2 * w2 + 5 * w1 + 11 * w0) / 6,
                -(w1 + 7 * w0) / 6, w0 / 3], axis=-1)
  u_all = np.stack([_roll(u, i) for i in (1, 0, -1, -2, -3)], axis=-1)
  return (c * u_all).sum(axis=-1)


# ----------------------------------------------------------------------------
# equations.py
# ----------------------------------------------------------------------------

# kind -> variant -> (derivative names, derivative orders); equations.py:230-587
DERIVATIVES = {
    ('burgers', 'plain'): (('u_x', 'u_xx'), (1, 2)),
    ('burgers', 'conservative'): (('u', 'u_x'), (0, 1)),
    ('burgers', 'godunov'): (('u_minus', 'u_plus', 'u_x'), (0, 0, 1)),
    ('kdv', 'plain'): (('u_x', 'u_xxx'), (1, 3)),
    ('kdv', 'conservative'): (('u', 'u_xx'), (0, 2)),
    ('kdv', 'godunov'): (('u_minus', 'u_plus', 'u_xx'), (0, 0, 2)),
    ('ks', 'plain'): (('u_x', 'u_xx', 'u_xxxx'), (1, 2, 4)),
    ('ks', 'conservative'): (('u', 'u_x', 'u_xxx'), (0, 1, 3)),
    ('ks', 'godunov'): (('u_minus', 'u_plus', 'u_x', 'u_xxx'), (0, 0, 1, 3)),
}
DEFAULT_PERIOD = {'burgers': 2 * np.pi, 'kdv': 32.0, 'ks': 64.0}    # :239,:385,:493
STANDARD_DEVIATION = {'burgers': 0.7917, 'kdv': 0.594, 'ks': 0.299}  # :264-267,:405-408,:510-513
TIME_STEP = {'burgers': 1e-3, 'kdv': 2.5e-5, 'ks': 2.5e-5}           # :259-262,:400-403,:505-508


def staggered_first_derivative(y, dx):
  """equations.py:305-320."""
  return (1 / dx) * (np.concatenate([y[..., 1:], y[..., :1]], axis=-1) - y)


def godunov_convective_flux(u_minus, u_plus):
  """equations.py:341-349."""
  m2, p2 = u_minus ** 2, u_plus ** 2
  return 0.5 * np.where(u_minus <= u_plus, np.minimum(m2, p2), np.maximum(m2, p2))


class RandomForcing(object):
  """equations.py:196-227 (RandomState draw order a, omega, k, phi: :207-212)."""

  def __init__(self, period, reference_num_points, resample_factor, resample_method,
               nparams=20, seed=0, amplitude=1, k_min=1, k_max=3):
    rs = np.random.RandomState(seed)
    self.a = 0.5 * amplitude * rs.uniform(-1, 1, size=(nparams, 1))
    self.omega = rs.uniform(-0.4, 0.4, size=(nparams, 1))
    k_values = np.arange(k_min, k_max + 1)
    self.k = rs.choice(np.concatenate([-k_values, k_values]), size=(nparams, 1))
    self.phi = rs.uniform(0, 2 * np.pi, size=(nparams, 1))
    self.period = period
    self.reference_x = period / reference_num_points * np.arange(reference_num_points)
    self.resample_factor = resample_factor
    self.resample_method = resample_method

  def __call__(self, t, dtype=np.float64):
    """dtype=float32 mimics the TF path, where t is a float32 placeholder
    (integrate.py:57) and every NumPy constant is converted to float32."""
    f = np.dtype(dtype).type
    spatial_phase = (2 * np.pi * self.k * self.reference_x / self.period).astype(dtype)
    arg = self.omega.astype(dtype) * f(t) + spatial_phase + self.phi.astype(dtype)
    ref = (self.a.astype(dtype) * np.sin(arg)).sum(axis=0)
    fn = resample_mean if self.resample_method == 'mean' else subsample
    return fn(ref, self.resample_factor).astype(dtype)


class EquationSpec(object):
  """The subset of equations.Equation (equations.py:71-193) the hot path needs."""

  def __init__(self, kind, variant='plain', num_points=256, resample_factor=1,
               period=None, random_seed=0, eta=0.04, k_min=1, k_max=3):
    self.kind, self.variant = kind, variant
    self.names, self.orders = DERIVATIVES[(kind, variant)]
    self.conservative = variant != 'plain'
    self.grid_offset = STAGGERED if self.conservative else CENTERED
    self.num_points = num_points
    self.resample_factor = resample_factor
    self.period = DEFAULT_PERIOD[kind] if period is None else period
    self.dx = self.period / num_points
    self.x = self.dx * np.arange(num_points)
    self.random_seed = random_seed
    self.eta = eta
    self.standard_deviation = STANDARD_DEVIATION[kind]
    self.time_step = TIME_STEP[kind]
    self.forcing = RandomForcing(
        self.period, num_points * resample_factor, resample_factor,
        'mean' if self.conservative else 'subsample',
        nparams=20 if kind == 'burgers' else 10, seed=random_seed,
        k_min=k_min, k_max=k_max)

  def initial_value(self):
    """equations.py:256-257 (zeros), :397-398 and :515-516 (forcing(0))."""
    if self.kind == 'burgers':
      return np.zeros(self.num_points)
    return self.forcing(0)

  def equation_of_motion(self, y, d):
    """equations.py:269-274,331-338,360-370,410-415,450-457,468-478,518-524,559-567,576-587."""
    kind, variant, dx = self.kind, self.variant, self.dx
    if variant == 'plain':
      if kind == 'burgers':
        return self.eta * d['u_xx'] - y * d['u_x']
      if kind == 'kdv':
        return -6 * y * d['u_x'] - d['u_xxx']
      return -y * d['u_x'] - d['u_xxxx'] - d['u_xx']
    if variant == 'conservative':
      u = d['u']
      if kind == 'burgers':
        flux = 0.5 * u ** 2 - self.eta * d['u_x']
      elif kind == 'kdv':
        flux = 3 * u ** 2 + d['u_xx']
      else:
        flux = 0.5 * u ** 2 + d['u_xxx'] + d['u_x']
    else:
      g = godunov_convective_flux(d['u_minus'], d['u_plus'])
      if kind == 'burgers':
        flux = g - self.eta * d['u_x']
      elif kind == 'kdv':
        flux = 6 * g + d['u_xx']
      else:
        flux = d['u_xxx'] + d['u_x'] + g
    return -staggered_first_derivative(flux, dx)

  def finalize_time_derivative(self, t, y_t, dtype=np.float64):
    """equations.py:276-277 (Burgers adds forcing) / :137-155 (identity)."""
    if self.kind == 'burgers':
      return y_t + self.forcing(t, dtype=dtype)
    return y_t


# ----------------------------------------------------------------------------
# model.py
# ----------------------------------------------------------------------------


class NetSpec(object):
  """The hparams that shape predict_coefficients (training.py:133-141 defaults)."""

  def __init__(self, num_layers=3, filter_size=32, kernel_size=5, nonlinearity='relu',
               polynomial_accuracy_order=1, polynomial_accuracy_scale=1.0,
               coefficient_grid_min_size=6, ensure_unbiased_coefficients=False,
               model_target='coefficients'):
    self.num_layers = num_layers
    self.filter_size = filter_size
    self.kernel_size = kernel_size
    self.nonlinearity = nonlinearity
    self.polynomial_accuracy_order = polynomial_accuracy_order
    self.polynomial_accuracy_scale = polynomial_accuracy_scale
    self.coefficient_grid_min_size = coefficient_grid_min_size
    self.ensure_unbiased_coefficients = ensure_unbiased_coefficients
    self.model_target = model_target


def coefficient_grid(eq, net):
  """model.py:445-448."""
  return regular_grid(eq.grid_offset, 0, net.coefficient_grid_min_size, eq.dx)


def accuracy_layers(eq, net):
  """model.py:478-490."""
  grid = coefficient_grid(eq, net)
  method = FINITE_VOLUMES if eq.conservative else FINITE_DIFFERENCES
  return [PolynomialAccuracyLayer(grid, method, order, net.polynomial_accuracy_order,
                                  out_scale=net.polynomial_accuracy_scale)
          for order in eq.orders]


def layer_shapes(eq, net):
  """Kernel shapes [(k, cin, cout), ...] of the conv stack (model.py:455-458,492-495)."""
  if net.model_target == 'space_derivatives':                         # model.py:571-576
    cout = len(eq.orders)
  elif net.model_target in ('time_derivative', 'flux'):               # model.py:603-615
    cout = 1
  elif net.polynomial_accuracy_order:
    cout = sum(l.input_size for l in accuracy_layers(eq, net))
  else:
    cout = len(eq.orders) * coefficient_grid(eq, net).size             # model.py:464-467
  shapes, cin = [], 1
  for _ in range(net.num_layers - 1):
    shapes.append((net.kernel_size, cin, net.filter_size))
    cin = net.filter_size
  if net.num_layers > 0:
    shapes.append((net.kernel_size, cin, cout))
  return shapes


def glorot_weights(eq, net, seed=0, last_layer_scale=1.0, bias_scale=0.0):
  """Deterministic stand-in for a trained checkpoint: Glorot-uniform kernels (the
  tf.layers.conv1d default initialiser), biases zero unless bias_scale>0."""
  rs = np.random.RandomState(seed)
  weights = []
  shapes = layer_shapes(eq, net)
  for i, (k, cin, cout) in enumerate(shapes):
    limit = math.sqrt(6.0 / (k * cin + k * cout))
    w = rs.uniform(-limit, limit, size=(k, cin, cout))
    if i == len(shapes) - 1:
      w = w * last_layer_scale
    b = bias_scale * rs.uniform(-1, 1, size=(cout,))
    weights.append((w.astype(np.float32), b.astype(np.float32)))
  return weights


def predict_coefficients(inputs, eq, net, weights, dtype=np.float32):
  """model.py:420-513.  inputs [b, x] -> [b, x, derivative, coefficient]."""
  inputs = np.asarray(inputs, dtype=dtype)
  if inputs.shape[-1] != eq.num_points:                                # model.py:53-56
    raise ValueError('solution has unexpected size for equation')
  x = inputs[:, :, None] / np.dtype(dtype).type(eq.standard_deviation)  # :450-451
  n_hidden = net.num_layers - 1
  for w, b in weights[:n_hidden]:
    x = conv1d_periodic_layer(x, w.astype(dtype), b.astype(dtype), net.nonlinearity)
  grid = coefficient_grid(eq, net)
  if not net.polynomial_accuracy_order:                                # :460-475
    w, b = weights[n_hidden]
    x = conv1d_periodic_layer(x, w.astype(dtype), b.astype(dtype), None)
    out = x.reshape(inputs.shape + (len(eq.orders), grid.size))
    if net.ensure_unbiased_coefficients:
      out = out - out.mean(axis=-1, keepdims=True)
    return out
  layers_ = accuracy_layers(eq, net)
  if net.num_layers > 0:
    w, b = weights[n_hidden]
    x = conv1d_periodic_layer(x, w.astype(dtype), b.astype(dtype), None)  # :492-495
  else:
    x = np.broadcast_to(weights[0].astype(dtype), inputs.shape + (weights[0].size,))  # :496-502
  outs, start = [], 0
  for layer in layers_:                                                # :504-511
    outs.append(layer.apply(x[..., start:start + layer.input_size]))
    start += layer.input_size
  return np.stack(outs, axis=-2)


def extract_patches(inputs, size):
  """model.py:516-533."""
  padded = pad_periodic(np.asarray(inputs)[..., None], size - 1, center=True)[..., 0]
  n = inputs.shape[-1]
  return np.stack([padded[:, i:i + n] for i in range(size)], axis=-1)


def apply_coefficients(coefs, inputs):
  """model.py:536-548: einsum('bxdi,bxi->bxd') on the UN-normalised inputs."""
  patches = extract_patches(inputs, coefs.shape[3])
  return np.einsum('bxdi,bxi->bxd', coefs, patches.astype(coefs.dtype))


def baseline_space_derivatives(inputs, eq, accuracy_order=1, dtype=np.float32):
  """model.py:59-112.  accuracy_order=None -> the 'exact' dispatch (:70-97)."""
  inputs = np.asarray(inputs, dtype=dtype)
  method = FINITE_VOLUMES if eq.conservative else FINITE_DIFFERENCES
  out = []
  for name, order in zip(eq.names, eq.orders):
    if accuracy_order is None and eq.variant == 'godunov' and eq.kind == 'burgers':
      if name == 'u_minus':
        d = _roll(weno_reconstruct_left(inputs), 1)                    # :83-84
      elif name == 'u_plus':
        d = _roll(weno_reconstruct_right(inputs), 1)                   # :86-87
      else:
        grid = regular_grid(eq.grid_offset, order, 3, eq.dx)           # :90-96
        d = reconstruct(inputs, grid, method, order)
    elif accuracy_order is None:
      d = spectral_derivative(inputs, order, eq.period)                # :78-80
    else:
      grid = regular_grid(eq.grid_offset, order, accuracy_order, eq.dx)  # :99-109
      d = reconstruct(inputs, grid, method, order)
    out.append(d.astype(dtype))
  return np.stack(out, axis=-1)


def apply_space_derivatives(derivs, inputs, eq):
  """model.py:115-135."""
  d = {name: derivs[..., i] for i, name in enumerate(eq.names)}
  return eq.equation_of_motion(inputs, d)


def multilayer_conv1d(inputs, eq, net, weights, dtype=np.float32):
  """model.py:551-568: the normalised conv stack with a linear last layer, [b, x] -> [b, x, targets]."""
  x = np.asarray(inputs, dtype=dtype)[:, :, None] / np.dtype(dtype).type(eq.standard_deviation)
  for i, (w, b) in enumerate(weights):
    act = net.nonlinearity if i < len(weights) - 1 else None
    x = conv1d_periodic_layer(x, w.astype(dtype), b.astype(dtype), act)
  return x


def predict_time_derivative(inputs, eq, net, weights, dtype=np.float32):
  """model.py:618-640, every model_target."""
  inputs = np.asarray(inputs, dtype=dtype)
  if net.model_target == 'time_derivative':                            # :603-606
    return multilayer_conv1d(inputs, eq, net, weights, dtype)[..., 0]
  if net.model_target == 'flux':                                       # :609-615 (note: no minus sign)
    return staggered_first_derivative(multilayer_conv1d(inputs, eq, net, weights, dtype)[..., 0],
                                      np.dtype(dtype).type(eq.dx)).astype(dtype)
  if net.model_target == 'space_derivatives':                          # :571-576
    derivs = multilayer_conv1d(inputs, eq, net, weights, dtype)
    return apply_space_derivatives(derivs, inputs, eq).astype(dtype)
  coefs = predict_coefficients(inputs, eq, net, weights, dtype=dtype)
  derivs = apply_coefficients(coefs, inputs)
  return apply_space_derivatives(derivs, inputs, eq).astype(dtype)


# ----------------------------------------------------------------------------
# integrate.py
# ----------------------------------------------------------------------------


class ModelDifferentiator(object):
  """integrate.py:48-71 (SavedModelDifferentiator), weights passed explicitly."""

  def __init__(self, eq, net, weights):
    self.eq, self.net, self.weights = eq, net, weights

  def __call__(self, t, y):
    y32 = np.asarray(y, dtype=np.float32)[None, :]
    y_t = predict_time_derivative(y32, self.eq, self.net, self.weights)[0]
    return self.eq.finalize_time_derivative(np.float32(t), y_t, dtype=np.float32)


class PolynomialDifferentiator(object):
  """integrate.py:74-105: float32 TF graph with fixed stencils."""

  def __init__(self, eq, accuracy_order=1):
    self.eq, self.accuracy_order = eq, accuracy_order

  def space_derivatives(self, y):
    y32 = np.asarray(y, dtype=np.float32)[None, :]
    d = baseline_space_derivatives(y32, self.eq, self.accuracy_order)[0]
    return {name: d[..., i] for i, name in enumerate(self.eq.names)}

  def __call__(self, t, y):
    y32 = np.asarray(y, dtype=np.float32)[None, :]
    d = baseline_space_derivatives(y32, self.eq, self.accuracy_order)
    y_t = apply_space_derivatives(d, y32, self.eq)[0].astype(np.float32)
    return self.eq.finalize_time_derivative(np.float32(t), y_t, dtype=np.float32)


class WENODifferentiator(object):
  """integrate.py:124-140: float64 WENO u-/u+, float32 4-point FV u_x (etc.)."""

  def __init__(self, eq, non_weno_accuracy_order=3, dtype=np.float64):
    if eq.variant != 'godunov':
      raise ValueError('invalid equation')
    self.eq = eq
    self.poly = PolynomialDifferentiator(eq, non_weno_accuracy_order)
    self.dtype = dtype

  def __call__(self, t, y):
    y = np.asarray(y, dtype=self.dtype)
    d = self.poly.space_derivatives(y)
    d['u_minus'] = np.roll(weno_reconstruct_left(y), 1)                # :137
    d['u_plus'] = np.roll(weno_reconstruct_right(y), 1)                # :138
    y_t = self.eq.equation_of_motion(y, d)
    return self.eq.finalize_time_derivative(t, y_t, dtype=self.dtype)


class SpectralDifferentiator(object):
  """integrate.py:108-121 (scipy.fftpack.diff == duckarray.spectral_derivative for even N)."""

  def __init__(self, eq):
    self.eq = eq

  def __call__(self, t, y):
    d = {n: spectral_derivative(y, o, self.eq.period) for n, o in zip(self.eq.names, self.eq.orders)}
    return self.eq.finalize_time_derivative(t, self.eq.equation_of_motion(y, d))


def odeint(y0, differentiator, times, method='RK23'):
  """integrate.py:143-169: SciPy adaptive RK23, max_step=0.01, NaN padding on failure."""
  import scipy.integrate
  sol = scipy.integrate.solve_ivp(differentiator, (times[0], times[-1]), y0,
                                  t_eval=times, max_step=0.01, method=method)
  y = sol.y.T
  missing = len(times) - y.shape[0]
  if missing:
    y = np.pad(y, ((0, missing), (0, 0)), mode='constant', constant_values=np.nan)
  return y, sol.nfev


# Bogacki-Shampine 3(2) tableau, as in scipy/integrate/_ivp/rk.py (RK23).
RK23_C = (0.0, 0.5, 0.75)
RK23_A = ((), (0.5,), (0.0, 0.75))
RK23_B = (2 / 9, 1 / 3, 4 / 9)
TABLEAUS = {
    'rk3': (RK23_C, RK23_A, RK23_B),
    'midpoint': ((0.0, 0.5), ((), (0.5,)), (0.0, 1.0)),   # tf.contrib.integrate.odeint_fixed, model.py:156-157
    'euler': ((0.0,), ((),), (1.0,)),
    'rk4': ((0.0, 0.5, 0.5, 1.0), ((), (0.5,), (0.0, 0.5), (0.0, 0.0, 1.0)),
            (1 / 6, 1 / 3, 1 / 3, 1 / 6)),
}


def fixed_step_integrate(rhs, y0, t0, dt, num_steps, save_every=1, scheme='rk3',
                         state_dtype=np.float64):
  """Fixed-step explicit RK with a float64 state and whatever dtype `rhs` returns
  (float32 for the TF-path differentiators) -- what SciPy's RK23 degenerates to when
  the controller is pinned at max_step (rk.py rk_step), without FSAL reuse.

  rhs(t, y[batch, x]) -> [batch, x].  Returns [num_saves, batch, x] (state_dtype).
  """
  c, a, b = TABLEAUS[scheme]
  y = np.array(y0, dtype=state_dtype)
  out = []
  for step in range(num_steps):
    t = t0 + step * dt
    ks = []
    for s in range(len(c)):
      ys = y
      if s:
        acc = np.zeros_like(y)
        for j, coeff in enumerate(a[s]):
          if coeff:
            acc = acc + coeff * ks[j]
        ys = y + dt * acc
      ks.append(np.asarray(rhs(t + c[s] * dt, ys), dtype=state_dtype))
    acc = np.zeros_like(y)
    for j, coeff in enumerate(b):
      if coeff:
        acc = acc + coeff * ks[j]
    y = y + dt * acc
    if (step + 1) % save_every == 0:
      out.append(y.copy())
  return np.stack(out) if out else np.zeros((0,) + y.shape, dtype=state_dtype)


def batched_rhs(eqs, net=None, weights=None, mode='learned', accuracy_order=1,
                weno_dtype=np.float32):
  """Batched RHS over per-sample equations (they differ only in forcing seeds).

  Returns f(t, y[batch, x]) -> [batch, x]: float32 for the TF-path modes, weno_dtype
  for 'weno'.
  """
  eq0 = eqs[0]

  def f(t, y):
    if mode == 'learned':
      y32 = np.asarray(y, dtype=np.float32)
      y_t = predict_time_derivative(y32, eq0, net, weights)
      dt_ = np.float32
    elif mode == 'fd':
      y32 = np.asarray(y, dtype=np.float32)
      d = baseline_space_derivatives(y32, eq0, accuracy_order)
      y_t = apply_space_derivatives(d, y32, eq0).astype(np.float32)
      dt_ = np.float32
    elif mode == 'weno':
      yw = np.asarray(y, dtype=weno_dtype)
      d32 = baseline_space_derivatives(yw.astype(np.float32), eq0, 3)
      d = {name: d32[..., i] for i, name in enumerate(eq0.names)}
      d['u_minus'] = _roll(weno_reconstruct_left(yw), 1)
      d['u_plus'] = _roll(weno_reconstruct_right(yw), 1)
      y_t = eq0.equation_of_motion(yw, d).astype(weno_dtype)
      dt_ = weno_dtype
    else:
      raise ValueError(mode)
    if eq0.kind == 'burgers':
      forcing = np.stack([e.forcing(dt_(t), dtype=dt_) for e in eqs])
      y_t = y_t + forcing
    return y_t

  return f
