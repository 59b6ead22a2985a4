"""RowSolver: a Python handle on one libddd1d configuration.

Everything numerical happens in the CUDA library; this module only translates the
reference's objects (Equation, hparams, TF-layout weights, NumPy stencil tables)
into ddd1d_* calls and owns the torch tensors used as device buffers.
"""
import ctypes
import json

import numpy as np

from . import _lib
from . import polynomials

_KIND = {'burgers': _lib.BURGERS, 'kdv': _lib.KDV, 'ks': _lib.KS}
_VARIANT = {'plain': _lib.PLAIN, 'conservative': _lib.CONSERVATIVE, 'godunov': _lib.GODUNOV}


def _torch():
  import torch
  if not torch.cuda.is_available():
    raise RuntimeError('ddd1d_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
  return torch


def _as_list(equations):
  return list(equations) if isinstance(equations, (list, tuple)) else [equations]


def _check_same_grid(equations):
  first = equations[0]
  for e in equations[1:]:
    if (type(e) is not type(first) or e.grid.solution_num_points != first.grid.solution_num_points
        or e.grid.period != first.grid.period or e.grid.resample_factor != first.grid.resample_factor):
      raise ValueError('all equations of a batch must share type and grid')
  return first


def method_for(equation):
  return (polynomials.Method.FINITE_VOLUMES if equation.CONSERVATIVE
          else polynomials.Method.FINITE_DIFFERENCES)


def baseline_windows(equation, accuracy_order):
  """Window-form standard stencils per derivative (model.py:99-109)."""
  rows = []
  for order in equation.DERIVATIVE_ORDERS:
    grid = polynomials.regular_grid(equation.GRID_OFFSET, order, accuracy_order,
                                    equation.grid.solution_dx)
    rows.append(_lib.to_window(polynomials.coefficients(grid, method_for(equation), order)))
  return np.stack(rows)


def coefficient_grid(equation, hparams):
  """model.py:445-448."""
  return polynomials.regular_grid(equation.GRID_OFFSET, derivative_order=0,
                                  accuracy_order=hparams.coefficient_grid_min_size,
                                  dx=equation.grid.solution_dx)


def accuracy_layers(equation, hparams):
  """model.py:478-490."""
  grid = coefficient_grid(equation, hparams)
  return [polynomials.PolynomialAccuracyLayer(
      grid=grid, method=method_for(equation), derivative_order=order,
      accuracy_order=hparams.polynomial_accuracy_order,
      out_scale=hparams.polynomial_accuracy_scale) for order in equation.DERIVATIVE_ORDERS]


def expected_layer_shapes(equation, hparams):
  """[(kernel_size, cin, cout)] of the conv stack (model.py:455-458,464-467,492-495)."""
  grid_size = coefficient_grid(equation, hparams).size
  target = getattr(hparams, 'model_target', 'coefficients')
  if target == 'space_derivatives':                       # model.py:571-576
    outputs = len(equation.DERIVATIVE_ORDERS)
  elif target in ('time_derivative', 'flux'):             # model.py:603-615
    outputs = 1
  elif target != 'coefficients':
    raise NotImplementedError('unrecognized model_target: {}'.format(target))
  elif hparams.polynomial_accuracy_order:
    outputs = sum(layer.input_size for layer in accuracy_layers(equation, hparams))
  else:
    outputs = len(equation.DERIVATIVE_ORDERS) * grid_size
  shapes, cin = [], 1
  for _ in range(hparams.num_layers - 1):
    shapes.append((hparams.kernel_size, cin, hparams.filter_size))
    cin = hparams.filter_size
  shapes.append((hparams.kernel_size, cin, outputs))
  return shapes


class RowSolver(object):
  """One ddd1d_handle plus the per-sample forcing it was given."""

  def __init__(self, equations, mode, hparams=None, weights=None, accuracy_order=1,
               weno_real='float32', device=None, forcing=True, engine='auto'):
    torch = _torch()
    self._lib = _lib.load()
    self.equations = _as_list(equations)
    eq = _check_same_grid(self.equations)
    self.equation = eq
    self.num_points = eq.grid.solution_num_points
    self.num_derivatives = len(eq.DERIVATIVE_ORDERS)
    self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    self.mode = mode
    self.stencil_size = 0
    self.constant_coefficients = None
    constant_windows = None
    if mode == _lib.MODE_LEARNED and hparams.num_layers == 0:
      # model.py:496-502: no net at all -- one learned vector through the polynomial-accuracy layers, i.e.
      # constant stencils.  They are formed here exactly as the float32 graph forms them (float32 tables,
      # one rounding per op, polynomials.py:275-277) and run on the fixed-stencil kernel.
      if not hparams.polynomial_accuracy_order:
        raise NotImplementedError                            # model.py:461-462
      if getattr(hparams, 'model_target', 'coefficients') != 'coefficients':
        raise NotImplementedError('num_layers=0 with model_target={!r}'.format(hparams.model_target))
      layers0 = accuracy_layers(eq, hparams)
      vector = np.asarray(weights[0] if isinstance(weights, (list, tuple)) else weights, dtype=np.float32).ravel()
      if vector.size != sum(l.input_size for l in layers0):
        raise ValueError('coefficients vector has {} entries, the model needs {}'.format(
            vector.size, sum(l.input_size for l in layers0)))
      rows, start = [], 0
      for l in layers0:
        z = vector[start:start + l.input_size]
        start += l.input_size
        product = (z.astype(np.float64) @ l.nullspace.astype(np.float32).astype(np.float64)).astype(np.float32)
        rows.append(l.bias.astype(np.float32) + product)
      self.constant_coefficients = np.stack(rows)            # [derivative, grid point] float32
      self.stencil_size = self.constant_coefficients.shape[1]
      constant_windows = np.ascontiguousarray(_lib.to_window(self.constant_coefficients.astype(np.float64)))
      mode = self.mode = _lib.MODE_STENCIL

    cfg = _lib.Config()
    cfg.struct_bytes = ctypes.sizeof(_lib.Config)
    cfg.device = self.device.index
    cfg.equation = _KIND[eq.KIND]
    cfg.variant = _VARIANT[eq.VARIANT]
    cfg.mode = mode
    cfg.num_points = self.num_points
    cfg.num_derivatives = self.num_derivatives
    cfg.weno_real = _lib.REAL_F64 if weno_real in ('float64', np.float64) else _lib.REAL_F32
    cfg.dx = eq.grid.solution_dx
    cfg.eta = getattr(eq, 'eta', 0.0)
    cfg.standard_deviation = eq.standard_deviation
    cfg.engine = _lib.ENGINES[engine]
    layers = None
    if mode == _lib.MODE_LEARNED:
      if hparams.nonlinearity not in _lib.ACTIVATIONS:
        raise KeyError(hparams.nonlinearity)
      shapes = expected_layer_shapes(eq, hparams)
      weights = [(np.ascontiguousarray(k, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32))
                 for k, b in weights]
      if [tuple(k.shape) for k, _ in weights] != shapes:
        raise ValueError('weights have shapes {}, the model needs {}'.format(
            [tuple(k.shape) for k, _ in weights], shapes))
      grid_size = coefficient_grid(eq, hparams).size
      self.stencil_size = grid_size
      cfg.num_layers = hparams.num_layers
      cfg.filter_size = hparams.filter_size
      cfg.kernel_size = hparams.kernel_size
      cfg.activation = _lib.ACTIVATIONS[hparams.nonlinearity]
      cfg.net_outputs = shapes[-1][2]
      cfg.stencil_size = grid_size
      target = getattr(hparams, 'model_target', 'coefficients')
      if target != 'coefficients':
        cfg.projection = {'space_derivatives': _lib.PROJ_DERIVATIVES,
                          'time_derivative': _lib.PROJ_TIME_DERIVATIVE, 'flux': _lib.PROJ_FLUX}[target]
        cfg.stencil_size = 1
      elif hparams.polynomial_accuracy_order:
        cfg.projection = _lib.PROJ_NULLSPACE
        layers = accuracy_layers(eq, hparams)
      elif hparams.ensure_unbiased_coefficients:
        if 0 in eq.DERIVATIVE_ORDERS:
          raise ValueError('ensure_unbiased not yet supported for 0th order spatial derivatives')
        cfg.projection = _lib.PROJ_RAW_UNBIASED
      else:
        cfg.projection = _lib.PROJ_RAW
    self._handle = ctypes.c_void_p()
    _lib.check(self._lib.ddd1d_create(ctypes.byref(cfg), ctypes.byref(self._handle)))

    if mode == _lib.MODE_LEARNED:
      for i, (kernel, bias) in enumerate(weights):
        self._check(self._lib.ddd1d_set_layer(self._handle, i, _lib.host_ptr(kernel), _lib.host_ptr(bias),
                                               kernel.shape[0], kernel.shape[1], kernel.shape[2]))
      if layers is not None:
        bias = np.ascontiguousarray(np.stack([l.bias_window() for l in layers]))
        basis = np.ascontiguousarray(np.concatenate([l.nullspace_window() for l in layers], axis=0))
        sizes = np.ascontiguousarray([l.input_size for l in layers], dtype=np.int32)
        self._check(self._lib.ddd1d_set_stencils(self._handle, _lib.host_ptr(bias)))
        self._check(self._lib.ddd1d_set_projection(self._handle, _lib.host_ptr(basis), _lib.host_ptr(sizes)))
    else:
      windows = (constant_windows if constant_windows is not None
                 else np.ascontiguousarray(baseline_windows(eq, accuracy_order)))
      self._check(self._lib.ddd1d_set_stencils(self._handle, _lib.host_ptr(windows)))

    self.forced = bool(forcing and eq.FORCED)
    if self.forced:
      stack = lambda name: np.ascontiguousarray(
          np.stack([getattr(e.forcing, name)[:, 0] for e in self.equations]).astype(np.float64))
      a, omega, k, phi = stack('a'), stack('omega'), stack('k'), stack('phi')
      self._check(self._lib.ddd1d_set_forcing(
          self._handle, _lib.host_ptr(a), _lib.host_ptr(omega), _lib.host_ptr(k), _lib.host_ptr(phi),
          a.shape[0], a.shape[1], eq.grid.resample_factor, int(eq.grid.resample_method == 'mean'),
          float(eq.grid.period)))

  # -- plumbing -----------------------------------------------------------------------
  def _check(self, code):
    _lib.check(code, self._handle)

  def close(self):
    if getattr(self, '_handle', None):
      self._lib.ddd1d_destroy(self._handle)
      self._handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # interpreter shutdown
      pass

  def _stream(self):
    return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

  def _rows(self, u, dtype=None):
    """[batch, N] contiguous CUDA tensor from a tensor or array."""
    torch = _torch()
    dtype = dtype or torch.float32
    t = torch.as_tensor(u)
    if t.dim() == 1:
      t = t[None]
    if t.dim() != 2 or t.shape[1] != self.num_points:
      raise ValueError('solution has unexpected size for equation: {} vs {}'.format(
          t.shape[-1], self.num_points))
    return t.to(device=self.device, dtype=dtype).contiguous()

  def _offset_ok(self, batch, sample_offset):
    if self.forced and sample_offset + batch > len(self.equations):
      raise ValueError('batch of {} at offset {} exceeds the {} equations given'.format(
          batch, sample_offset, len(self.equations)))

  # -- operations ---------------------------------------------------------------------
  def rhs(self, t, u, sample_offset=0):
    """dy/dt [batch, N] (float32 in/out, or float64 in/out) on the device."""
    torch = _torch()
    src = torch.as_tensor(u)
    dtype = torch.float64 if src.dtype == torch.float64 else torch.float32
    rows = self._rows(src, dtype)
    self._offset_ok(rows.shape[0], sample_offset)
    out = torch.empty_like(rows)
    fn = self._lib.ddd1d_rhs_f64 if dtype == torch.float64 else self._lib.ddd1d_rhs
    self._check(fn(self._handle, float(t), rows.data_ptr(), out.data_ptr(), rows.shape[0],
                   sample_offset, self._stream()))
    return out

  def rhs_host(self, t, y, sample_offset=0):
    """NumPy float64 [batch, N] -> float64, copies inside the library (SciPy's path)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    rows = y.reshape(-1, self.num_points)
    self._offset_ok(rows.shape[0], sample_offset)
    out = np.empty_like(rows)
    self._check(self._lib.ddd1d_rhs_host(self._handle, float(t), _lib.host_ptr(rows), _lib.host_ptr(out),
                                         rows.shape[0], sample_offset))
    return out.reshape(y.shape)

  def coefficients(self, u):
    torch = _torch()
    rows = self._rows(u)
    if self.constant_coefficients is not None:               # num_layers = 0: tf.tile of one vector
      const = torch.as_tensor(self.constant_coefficients, device=self.device)
      return const.expand(rows.shape + const.shape).contiguous()
    out = torch.empty(rows.shape + (self.num_derivatives, self.stencil_size), device=self.device,
                      dtype=torch.float32)
    self._check(self._lib.ddd1d_coefficients(self._handle, rows.data_ptr(), out.data_ptr(), rows.shape[0],
                                             self._stream()))
    return out

  def space_derivatives(self, u):
    torch = _torch()
    rows = self._rows(u)
    out = torch.empty(rows.shape + (self.num_derivatives,), device=self.device, dtype=torch.float32)
    self._check(self._lib.ddd1d_space_derivatives(self._handle, rows.data_ptr(), out.data_ptr(),
                                                  rows.shape[0], self._stream()))
    return out

  def integrate(self, u0, t0, dt, num_steps, save_every=1, scheme='rk3', sample_offset=0,
                return_first_bad=False, float32_state=False):
    """Fused fixed-step integration; returns snapshots [num_steps // save_every, batch, N].
    float32_state=True rounds the carried solution to float32 after every step (the reference's TF-graph
    unroll, model.py:138-159); the default carries it in float64 like SciPy (integrate.py:154)."""
    torch = _torch()
    rows = self._rows(u0)
    self._offset_ok(rows.shape[0], sample_offset)
    nsave = num_steps // save_every
    snaps = torch.empty((nsave,) + tuple(rows.shape), device=self.device, dtype=torch.float32)
    bad = torch.empty(rows.shape[0], device=self.device, dtype=torch.int32)
    self._check(self._lib.ddd1d_integrate(
        self._handle, float(t0), float(dt), int(num_steps), int(save_every),
        _lib.SCHEMES[scheme] | (_lib.STATE_F32 if float32_state else 0),
        rows.data_ptr(), snaps.data_ptr(), bad.data_ptr(), rows.shape[0], sample_offset, self._stream()))
    return (snaps, bad) if return_first_bad else snaps

  def odeint(self, u0, times, rtol=1e-3, atol=1e-6, max_step=0.01, sample_offset=0):
    """SciPy-RK23 twin on the device for every row (ddd1d_integrate_adaptive).
    Returns (y float64 [len(times), batch, N] with NaN padding, nfev int32 [batch],
    status int32 [batch])."""
    torch = _torch()
    times = np.ascontiguousarray(times, dtype=np.float64)
    src = torch.as_tensor(u0)
    f64 = src.dtype == torch.float64
    rows = self._rows(src, torch.float64 if f64 else torch.float32)
    self._offset_ok(rows.shape[0], sample_offset)
    y = torch.empty((len(times),) + tuple(rows.shape), device=self.device, dtype=torch.float64)
    nfev = torch.zeros(rows.shape[0], device=self.device, dtype=torch.int32)
    status = torch.zeros(rows.shape[0], device=self.device, dtype=torch.int32)
    self._check(self._lib.ddd1d_integrate_adaptive(
        self._handle, _lib.host_ptr(times), len(times), float(rtol), float(atol), float(max_step),
        None if f64 else rows.data_ptr(), rows.data_ptr() if f64 else None, y.data_ptr(),
        nfev.data_ptr(), status.data_ptr(), rows.shape[0], sample_offset, self._stream()))
    return y, nfev, status

  def integrate_host(self, u0, t0, dt, num_steps, save_every=1, scheme='rk3', sample_offset=0):
    """Host buffers in and out through ddd1d_integrate_host (copies inside the library)."""
    u0 = np.ascontiguousarray(u0, dtype=np.float32).reshape(-1, self.num_points)
    self._offset_ok(u0.shape[0], sample_offset)
    nsave = num_steps // save_every
    snaps = np.empty((nsave,) + u0.shape, dtype=np.float32)
    bad = np.empty(u0.shape[0], dtype=np.int32)
    self._check(self._lib.ddd1d_integrate_host(
        self._handle, float(t0), float(dt), int(num_steps), int(save_every), _lib.SCHEMES[scheme],
        _lib.host_ptr(u0), _lib.host_ptr(snaps), _lib.host_ptr(bad), u0.shape[0], sample_offset))
    return snaps, bad

  def engine(self):
    """'ffma', 'tensor', 'tensor_f16x2' or 'tensor_f16': the kernel the next launch will use."""
    code = self._lib.ddd1d_engine(self._handle)
    if code < 0:
      self._check(code)
    return _lib.ENGINE_NAMES[code]

  def launch_count(self):
    return int(self._lib.ddd1d_launch_count(self._handle))

  def launch_shape(self, batch):
    grid, block, smem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    self._check(self._lib.ddd1d_launch_shape(self._handle, int(batch), ctypes.byref(grid),
                                             ctypes.byref(block), ctypes.byref(smem)))
    return dict(grid=grid.value, block=block.value, shared_bytes=smem.value)


def stencil_solver(equations, accuracy_order=1, **kwargs):
  return RowSolver(equations, _lib.MODE_STENCIL, accuracy_order=accuracy_order, **kwargs)


def weno_solver(equations, non_weno_accuracy_order=3, **kwargs):
  return RowSolver(equations, _lib.MODE_WENO, accuracy_order=non_weno_accuracy_order, **kwargs)


def learned_solver(equations, hparams, weights, **kwargs):
  return RowSolver(equations, _lib.MODE_LEARNED, hparams=hparams, weights=weights, **kwargs)


def weights_fingerprint(hparams, weights):
  import hashlib
  h = hashlib.sha1(json.dumps(hparams.values(), sort_keys=True, default=str).encode())
  for item in weights:
    for array in (item if isinstance(item, (list, tuple)) else (item,)):   # (kernel, bias) | num_layers=0 vector
      h.update(np.ascontiguousarray(array, dtype=np.float32).tobytes())
  return h.hexdigest()
