"""Model-level entry points -- mirror of the inference half of
pde_superresolution/model.py (lines 42-159 and 411-661).

The reference finds its conv weights through TensorFlow variable scopes
(implicit global state restored by tf.train.Saver); here they are an explicit
argument: ``weights = [(kernel[k, cin, cout], bias[cout]), ...]`` in TF layout,
one pair per conv layer in creation order.  All tensors are torch CUDA tensors
(NumPy inputs are accepted and moved); every function runs the CUDA library.
"""
import numpy as np

from . import equations as equations_lib
from . import polynomials
from . import runtime

FINITE_DIFF = polynomials.Method.FINITE_DIFFERENCES
FINITE_VOL = polynomials.Method.FINITE_VOLUMES

import collections
import os

_SOLVERS = collections.OrderedDict()     # least recently used first
_MAX_SOLVERS = 32
_FINGERPRINTS = {}                       # id(weights) -> (weights, fingerprint): hash each weight list once


def _cached(key, build):
  """Solver cache with least-recently-used eviction (one handle at a time: callers never see a
  solver they hold disappear wholesale)."""
  if key in _SOLVERS:
    _SOLVERS.move_to_end(key)
    return _SOLVERS[key]
  while len(_SOLVERS) >= _MAX_SOLVERS:
    _SOLVERS.popitem(last=False)
  _SOLVERS[key] = build()
  return _SOLVERS[key]


def _fingerprint(hparams, weights):
  """SHA-1 of (hparams, weights), computed once per weights object: func(y, t) evaluations call
  predict_* thousands of times with the same list."""
  hit = _FINGERPRINTS.get(id(weights))
  values = tuple(sorted((k, str(v)) for k, v in hparams.values().items()))
  if hit is not None and hit[0] is weights and hit[1] == values:
    return hit[2]
  fp = runtime.weights_fingerprint(hparams, weights)
  if len(_FINGERPRINTS) > 4 * _MAX_SOLVERS:
    _FINGERPRINTS.clear()
  _FINGERPRINTS[id(weights)] = (weights, values, fp)
  return fp


def assert_consistent_solution(equation, solution):
  """model.py:42-56."""
  if equation.grid.solution_num_points != solution.shape[-1]:
    raise ValueError('solution has unexpected size for equation: {} vs {}'.format(
        solution.shape[-1], equation.grid.solution_num_points))


def _learned(hparams, weights):
  if weights is None:
    raise ValueError('weights must be given explicitly: [(kernel[k,cin,cout], bias[cout]), ...]')
  _, equation = equations_lib.from_hparams(hparams)
  # (the engine override of the environment is part of the key: a cached solver keeps the engine it was built for)
  key = ('learned', _fingerprint(hparams, weights), os.environ.get('DDD1D_ENGINE', ''))
  return equation, _cached(key, lambda: runtime.learned_solver(equation, hparams, weights, forcing=False))


def predict_coefficients(inputs, hparams, weights=None, reuse=None):
  """[batch, x] -> [batch, x, derivative, coefficient] (model.py:420-513)."""
  del reuse
  equation, solver = _learned(hparams, weights)
  assert_consistent_solution(equation, inputs)
  return solver.coefficients(inputs)


def extract_patches(inputs, size):
  """[batch, x] -> [batch, x, size] periodic patches, ceil((size-1)/2) points to the
  left (model.py:516-533)."""
  import torch
  x = torch.as_tensor(inputs)
  left = -(-(size - 1) // 2)
  return torch.stack([torch.roll(x, left - i, dims=-1) for i in range(size)], dim=-1)


def apply_coefficients(coefficients, inputs):
  """einsum('bxdi,bxi->bxd') (model.py:536-548).  Stand-alone form for callers that
  hold coefficients; inside the integrator this contraction is fused."""
  import torch
  coefficients = torch.as_tensor(coefficients)
  patches = extract_patches(torch.as_tensor(inputs, device=coefficients.device), coefficients.shape[3])
  return torch.einsum('bxdi,bxi->bxd', coefficients, patches.to(coefficients.dtype))


def predict_space_derivatives(inputs, hparams, weights=None, reuse=None):
  """[batch, x] -> [batch, x, derivative] (model.py:579-600)."""
  del reuse
  if hparams.model_target not in ('coefficients', 'space_derivatives'):
    raise NotImplementedError('unrecognized model_target: {}'.format(hparams.model_target))   # model.py:599-600
  equation, solver = _learned(hparams, weights)
  assert_consistent_solution(equation, inputs)
  return solver.space_derivatives(inputs)


def predict_time_derivative(inputs, hparams, weights=None, reuse=None):
  """[batch, x] -> [batch, x], the equation of motion applied to the predicted
  derivatives, WITHOUT finalize_time_derivative (model.py:618-640)."""
  del reuse
  equation, solver = _learned(hparams, weights)      # every model_target (model.py:632-640)
  assert_consistent_solution(equation, inputs)
  return solver.rhs(0.0, inputs)


def apply_fixed_stencils(inputs, stencils):
  """Periodic centred application of constant stencils (up to 2), [batch, x] ->
  [batch, x, len(stencils)]: the device form of layers.nn_conv1d_periodic(center=True)
  used by polynomials.reconstruct."""
  import torch
  from . import _lib
  x = torch.as_tensor(inputs)
  n = x.shape[-1]
  if not 1 <= len(stencils) <= 2:
    raise ValueError('1 or 2 stencils at a time')
  carrier = equations_lib.BurgersEquation(n, period=float(n))     # D = 2 channels, dx = 1
  key = ('stencils', n, tuple(tuple(float(v) for v in np.asarray(s, dtype=np.float64).ravel()) for s in stencils))
  def build():
    solver = runtime.stencil_solver(carrier, forcing=False)
    rows = np.zeros((2, _lib.WINDOW))
    for i, s in enumerate(stencils):
      rows[i] = _lib.to_window(s)
    solver._check(solver._lib.ddd1d_set_stencils(solver._handle, _lib.host_ptr(np.ascontiguousarray(rows))))
    return solver
  return _cached(key, build).space_derivatives(x)[..., :len(stencils)]


def spectral_derivative(inputs, order, period):
  """duckarray.spectral_derivative (duckarray.py:105-112) with torch.fft (cuFFT)."""
  import torch
  x = torch.as_tensor(inputs)
  n = x.shape[-1]
  if n % 2:
    raise ValueError('spectral derivative only works for even length data')
  k = torch.fft.rfftfreq(n, d=1.0 / n, device=x.device)
  factor = (2j * np.pi / period * k.to(torch.complex128 if x.dtype == torch.float64 else torch.complex64)) ** order
  return torch.fft.irfft(factor * torch.fft.rfft(x), n=n)


def baseline_space_derivatives(inputs, equation, accuracy_order=None):
  """[batch, x] -> [batch, x, derivative] with standard stencils (explicit
  accuracy_order, model.py:99-109) or the equation's "exact" method (None,
  model.py:70-97: WENO for Godunov Burgers, spectral for KdV / KS)."""
  import torch
  assert_consistent_solution(equation, inputs)
  grid = equation.grid
  tag = (type(equation).__name__, grid.solution_num_points, grid.period, getattr(equation, 'eta', None))
  if accuracy_order is not None:
    solver = _cached(('fd', accuracy_order) + tag,
                     lambda: runtime.stencil_solver(equation, accuracy_order, forcing=False))
    return solver.space_derivatives(inputs)
  assert equation.exact_type() is type(equation)
  if equation.EXACT_METHOD is equations_lib.ExactMethod.WENO:
    solver = _cached(('weno',) + tag, lambda: runtime.weno_solver(equation, forcing=False))
    return solver.space_derivatives(inputs)
  if equation.EXACT_METHOD is equations_lib.ExactMethod.SPECTRAL:
    x = torch.as_tensor(inputs)
    x = x.cuda() if not x.is_cuda else x
    return torch.stack([spectral_derivative(x, order, grid.period)
                        for order in equation.DERIVATIVE_ORDERS], dim=-1)
  # ExactMethod.POLYNOMIAL: 6-point stencils (model.py:73-77)
  six = (0.5 + np.arange(-3, 3)) * grid.solution_dx
  stencils = [polynomials.coefficients(six, runtime.method_for(equation), order)
              for order in equation.DERIVATIVE_ORDERS]
  return torch.cat([apply_fixed_stencils(inputs, stencils[i:i + 2]) for i in range(0, len(stencils), 2)],
                   dim=-1)


def apply_space_derivatives(derivatives, inputs, equation):
  """[batch, x, derivative] + [batch, x] -> dy/dt via equation.equation_of_motion
  (model.py:115-135)."""
  named = {name: derivatives[..., i] for i, name in enumerate(equation.DERIVATIVE_NAMES)}
  return equation.equation_of_motion(inputs, named)


def integrate_ode(func, inputs, num_time_steps, time_step):
  """Fixed-step midpoint rule for an arbitrary func(y, t) on device tensors,
  [batch, x] -> [batch, x, num_time_steps] (model.py:138-159).  For the learned
  model prefer predict_time_evolution, which fuses all steps into one kernel."""
  import torch
  y = torch.as_tensor(inputs)
  out = []
  for step in range(num_time_steps):
    t = step * time_step
    k1 = func(y, t)
    y = y + time_step * func(y + (time_step / 2) * k1, t + time_step / 2)
    out.append(y)
  return torch.stack(out, dim=-1)


def predict_time_evolution(inputs, hparams, weights=None):
  """[batch, x] -> [batch, x, num_time_steps] with the learned model and the midpoint
  rule (model.py:643-661), all steps in one persistent kernel launch."""
  equation, solver = _learned(hparams, weights)
  assert_consistent_solution(equation, inputs)
  # float32 carry between steps, as tf.contrib.integrate.odeint_fixed on the float32 graph (model.py:156-157)
  snaps = solver.integrate(inputs, 0.0, equation.time_step, hparams.num_time_steps, 1, 'midpoint', float32_state=True)
  return snaps.permute(1, 2, 0)


# ---------------------------------------------------------------------------------
# direct-prediction variants and the stacked-result helpers (model.py:162-290, 551-615, 664-696)
# ---------------------------------------------------------------------------------
def _with_target(hparams, target):
  import copy
  hp = copy.deepcopy(hparams)
  hp.model_target = target
  return hp


def predict_space_derivatives_directly(inputs, hparams, weights=None, reuse=None):
  """model.py:571-576: the net's channels are the space derivatives, [batch, x, derivative]."""
  return predict_space_derivatives(inputs, _with_target(hparams, 'space_derivatives'), weights, reuse)


def predict_time_derivative_directly(inputs, hparams, weights=None, reuse=None):
  """model.py:603-606: the net's single channel is dy/dt, [batch, x]."""
  return predict_time_derivative(inputs, _with_target(hparams, 'time_derivative'), weights, reuse)


def predict_flux_directly(inputs, hparams, weights=None, reuse=None):
  """model.py:609-615: staggered_first_derivative of the net's single channel (no minus sign)."""
  return predict_time_derivative(inputs, _with_target(hparams, 'flux'), weights, reuse)


def result_stack(space_derivatives, time_derivative, integrated_solution=None):
  """[..., derivative] + [...] (+ [..., time]) -> [..., derivative + 1 (+ time)] (model.py:186-205)."""
  import torch
  tensors = [torch.as_tensor(space_derivatives), torch.as_tensor(time_derivative)[..., None]]
  if integrated_solution is not None:
    tensors.append(torch.as_tensor(integrated_solution))
  return torch.cat(tensors, dim=-1)


def result_unstack(tensor, equation):
  """Inverse of result_stack (model.py:208-234)."""
  d = len(equation.DERIVATIVE_ORDERS)
  integrated = tensor[..., d + 1:] if tensor.shape[-1] > d + 1 else None
  return tensor[..., :d], tensor[..., d], integrated


def baseline_time_evolution(inputs, num_time_steps, equation):
  """Midpoint-rule evolution with first-order-accurate standard stencils, all steps in one
  fused launch, [batch, x] -> [batch, x, num_time_steps] (model.py:162-183)."""
  assert_consistent_solution(equation, inputs)
  grid = equation.grid
  tag = (type(equation).__name__, grid.solution_num_points, grid.period, getattr(equation, 'eta', None))
  solver = _cached(('fd', 1) + tag, lambda: runtime.stencil_solver(equation, 1, forcing=False))
  snaps = solver.integrate(inputs, 0.0, equation.time_step, num_time_steps, 1, 'midpoint', float32_state=True)
  return snaps.permute(1, 2, 0)


def baseline_result(inputs, equation, num_time_steps=0, accuracy_order=None):
  """Space derivatives, time derivative and (optionally) the evolved solution of the baseline
  model, stacked on the last axis (model.py:245-275)."""
  if accuracy_order is None:
    equation = equation.to_exact()
  elif type(equation) in equations_lib.FLUX_EQUATION_TYPES:
    # (as in the reference, model.py:264-265: the registry is keyed by NAME, so this membership test on a
    #  class never holds and Godunov equations keep their own derivative set)
    equation = equation.to_conservative()
  space_derivatives = baseline_space_derivatives(inputs, equation, accuracy_order=accuracy_order)
  import torch
  rows = torch.as_tensor(inputs, device=space_derivatives.device).to(space_derivatives.dtype)
  time_derivative = apply_space_derivatives(space_derivatives, rows, equation)
  integrated = baseline_time_evolution(inputs, num_time_steps, equation) if num_time_steps else None
  return result_stack(space_derivatives, time_derivative, integrated)


def predict_result(inputs, hparams, weights=None):
  """The learned model's counterpart of baseline_result (model.py:664-696)."""
  import torch
  _, equation = equations_lib.from_hparams(hparams)
  if hparams.model_target in ('flux', 'time_derivative'):
    if hparams.space_derivatives_weight:
      raise ValueError('space derivatives are not predicted by model {}'.format(hparams.model_target))
    time_derivative = predict_time_derivative(inputs, hparams, weights)
    space_derivatives = torch.zeros(tuple(time_derivative.shape) + (len(equation.DERIVATIVE_ORDERS),),
                                    device=time_derivative.device, dtype=time_derivative.dtype)
  else:
    space_derivatives = predict_space_derivatives(inputs, hparams, weights)
    rows = torch.as_tensor(inputs, device=space_derivatives.device).to(space_derivatives.dtype)
    time_derivative = apply_space_derivatives(space_derivatives, rows, equation)
  integrated = predict_time_evolution(inputs, hparams, weights) if hparams.num_time_steps else None
  return result_stack(space_derivatives, time_derivative, integrated)
