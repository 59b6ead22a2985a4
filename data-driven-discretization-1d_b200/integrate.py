"""Time integration -- mirror of pde_superresolution/integrate.py.

Two ways in:
  * the reference's own surface: Differentiator objects called by SciPy's adaptive
    RK23 through odeint()/integrate() one sample at a time (integrate.py:143-279).
    Here every Differentiator.__call__ is one ddd1d_rhs launch instead of a
    sess.run, so existing callers work unchanged;
  * BatchIntegrator: the whole batch advanced by the fused persistent kernel with
    a fixed step (what RK23 degenerates to once its controller is pinned at
    max_step=0.01, integrate.py:154-155), rows resident on chip between snapshots.
"""
import functools
import logging

import numpy as np

from . import duckarray
from . import equations as equations_lib
from . import model
from . import runtime
from . import training

_DEFAULT_TIMES = np.linspace(0, 10, num=201)


# ---------------------------------------------------------------------------------
# result container (xarray is optional: the reference returns xarray.Dataset)
# ---------------------------------------------------------------------------------
class _Variable(object):
  def __init__(self, dims, data):
    self.dims, self.data = tuple(dims), np.asarray(data)

  @property
  def values(self):
    return self.data

  def __array__(self, dtype=None):
    return self.data if dtype is None else self.data.astype(dtype)


class Dataset(object):
  """Tiny stand-in with the slice of xarray.Dataset's interface the reference's
  callers use (ds['y'].data / .dims, ds.coords[...], ds.dims)."""

  def __init__(self, data_vars, coords):
    self.data_vars = {k: _Variable(*v) for k, v in data_vars.items()}
    self.coords = {k: np.asarray(v) for k, v in coords.items()}

  def __getitem__(self, name):
    if name in self.data_vars:
      return self.data_vars[name]
    return _Variable((name,) if self.coords[name].ndim else (), self.coords[name])

  @property
  def dims(self):
    out = {}
    for var in self.data_vars.values():
      out.update(zip(var.dims, var.data.shape))
    return out


def make_dataset(data_vars, coords):
  try:
    import xarray
    return xarray.Dataset(data_vars, coords=coords)
  except ImportError:
    return Dataset(data_vars, coords)


# ---------------------------------------------------------------------------------
# Differentiators
# ---------------------------------------------------------------------------------
class Differentiator(object):
  """Callable (t, y[N]) -> dy/dt[N] in NumPy float64 (integrate.py:40-45)."""

  def __call__(self, t, y):
    raise NotImplementedError


def _load_weights(source):
  if isinstance(source, (list, tuple)):
    return list(source)
  if isinstance(source, str) and source.endswith('.npz'):
    with np.load(source) as f:
      out, i = [], 0
      while 'kernel%d' % i in f.files:
        out.append((f['kernel%d' % i], f['bias%d' % i]))
        i += 1
    return out
  # a checkpoint directory (or prefix) written by tf.train.Saver (integrate.py:66-68)
  from . import checkpoint
  return checkpoint.load_conv_weights(source)


class SavedModelDifferentiator(Differentiator):
  """Learned-coefficient model (integrate.py:48-71).  `checkpoint_dir` may be the
  weights themselves, see _load_weights."""

  def __init__(self, checkpoint_dir, equation, hparams):
    self.equation = equation
    self.solver = runtime.learned_solver(equation, hparams, _load_weights(checkpoint_dir))

  def __call__(self, t, y):
    return self.solver.rhs_host(t, y)


class PolynomialDifferentiator(Differentiator):
  """Standard finite differences / volumes (integrate.py:74-105)."""

  def __init__(self, equation, accuracy_order=1):
    self.equation = equation
    self.accuracy_order = accuracy_order
    self._spectral = None
    if accuracy_order is not None:
      self.solver = runtime.stencil_solver(equation, accuracy_order)
    elif equation.EXACT_METHOD is equations_lib.ExactMethod.WENO:
      assert equation.exact_type() is type(equation)
      self.solver = runtime.weno_solver(equation)      # float32 WENO, model.py:81-97
    else:
      self.solver = None
      self._spectral = SpectralDifferentiator(equation)

  def __call__(self, t, y):
    if self.solver is None:
      return self._spectral(t, y)
    return self.solver.rhs_host(t, y)

  def calculate_space_derivatives(self, y):
    values = model.baseline_space_derivatives(
        np.asarray(y, dtype=np.float32)[None], self.equation, self.accuracy_order)[0].cpu().numpy()
    return {name: values[:, i] for i, name in enumerate(self.equation.DERIVATIVE_NAMES)}


class SpectralDifferentiator(Differentiator):
  """Fourier derivatives in float64 (integrate.py:108-121), via cuFFT."""

  def __init__(self, equation):
    self.equation = equation

  def __call__(self, t, y):
    import torch
    eq = self.equation
    u = torch.as_tensor(np.asarray(y, dtype=np.float64), device='cuda')
    derivs = {name: model.spectral_derivative(u, order, eq.grid.period)
              for name, order in zip(eq.DERIVATIVE_NAMES, eq.DERIVATIVE_ORDERS)}
    y_t = eq.equation_of_motion(u, derivs).cpu().numpy()
    return eq.finalize_time_derivative(t, y_t)


class WENODifferentiator(Differentiator):
  """5th-order WENO for Godunov-flux equations (integrate.py:124-140).  Like the reference, the
  reconstruction, flux and forcing are float64 and only the non-WENO stencil is float32;
  weno_real='float32' selects the all-float32 kernel (what model.baseline_space_derivatives does
  in TF, model.py:81-97)."""

  def __init__(self, equation, non_weno_accuracy_order=3, weno_real='float64'):
    if equation.VARIANT != 'godunov':
      raise ValueError('invalid equation: {}'.format(equation))
    self.equation = equation
    self.solver = runtime.weno_solver(equation, non_weno_accuracy_order, weno_real=weno_real)

  def __call__(self, t, y):
    return self.solver.rhs_host(t, y)


# ---------------------------------------------------------------------------------
# SciPy-driven integration, the reference's surface
# ---------------------------------------------------------------------------------
def odeint(y0, differentiator, times, method='RK23'):
  """scipy solve_ivp with max_step=0.01; diverged runs are NaN-padded, not raised
  (integrate.py:143-169).  Returns (y[time, x], nfev)."""
  import scipy.integrate
  sol = scipy.integrate.solve_ivp(differentiator, (times[0], times[-1]), y0, t_eval=times,
                                  max_step=0.01, method=method)
  y = sol.y.T
  logging.info('nfev: %r, status: %r, message: %s', sol.nfev, sol.status, sol.message)
  missing = len(times) - y.shape[0]
  if missing:
    y = np.pad(y, ((0, missing), (0, 0)), mode='constant', constant_values=np.nan)
  return y, sol.nfev


def smoothing_filter(x, alpha=-np.log(1e-15), order=2):
  """Exponential spectral low-pass (duckarray.py:115-128), cuFFT."""
  import torch
  t = torch.as_tensor(np.asarray(x, dtype=np.float64), device='cuda')
  n = t.shape[-1]
  if n % 2:
    raise ValueError('smoothing filter only works for even length data')
  eta = torch.arange(n // 2 + 1, device='cuda', dtype=torch.float64) / (n // 2)
  sigma = torch.exp(-alpha * eta ** (2 * order))
  return torch.fft.irfft(sigma * torch.fft.rfft(t), n=n).cpu().numpy()


def odeint_with_periodic_filtering(y0, differentiator, times, filter_interval, filter_order,
                                   method='RK23'):
  """Integrate in chunks of `filter_interval`, low-pass filtering the state between
  chunks and the whole record at the end (integrate.py:172-212)."""
  eps = 1e-8
  split_times = np.arange(times[0], times[-1] + eps, filter_interval)
  if not np.isin(split_times, times).all():
    raise ValueError('all times in filter_interval must be sampled')
  splits = np.searchsorted(times, split_times, side='right')
  pieces = [y0[np.newaxis, ...]]
  num_evals = 0
  for start, stop in zip(splits[:-1], splits[1:]):
    y, n = odeint(y0, differentiator, times[start - 1:stop], method=method)
    pieces.append(y[1:])
    y0 = smoothing_filter(y[-1], order=filter_order)
    num_evals += n
  y = np.concatenate(pieces, axis=0)
  assert y.shape == (times.size, y0.size)
  return smoothing_filter(y, order=filter_order), num_evals


def exact_differentiator(equation):
  """integrate.py:215-235."""
  if type(equation.to_exact()) is not type(equation):
    raise TypeError('an exact equation must be provided')
  if equation.EXACT_METHOD is equations_lib.ExactMethod.POLYNOMIAL:
    return PolynomialDifferentiator(equation, accuracy_order=None)
  if equation.EXACT_METHOD is equations_lib.ExactMethod.SPECTRAL:
    return SpectralDifferentiator(equation)
  if equation.EXACT_METHOD is equations_lib.ExactMethod.WENO:
    return WENODifferentiator(equation)
  raise TypeError('unexpected equation: {}'.format(equation))


def integrate(equation, differentiator, times=_DEFAULT_TIMES, warmup=0, integrate_method='RK23',
              filter_interval=None, filter_all_times=False):
  """Optional exact warm-up, then odeint; result {y: (time, x)} with coords time, x,
  num_evals (integrate.py:238-279)."""
  if filter_interval is not None:
    warmup_odeint = functools.partial(
        odeint_with_periodic_filtering, filter_interval=filter_interval,
        filter_order=max(equation.to_exact().DERIVATIVE_ORDERS))
  else:
    warmup_odeint = odeint
  if warmup:
    exact = equation.to_exact()
    if filter_interval is not None:
      warmup_times = np.arange(0, warmup + 1e-8, filter_interval)
    else:
      warmup_times = np.array([0, warmup])
    spun_up, _ = warmup_odeint(exact.initial_value(), exact_differentiator(exact),
                               times=warmup_times, method=integrate_method)
    y0 = equation.grid.resample(spun_up[-1, :])
  else:
    y0 = equation.initial_value()
  solve = warmup_odeint if filter_all_times else odeint
  solution, num_evals = solve(y0, differentiator, times=warmup + times, method=integrate_method)
  return make_dataset({'y': (('time', 'x'), solution)},
                      {'time': warmup + times, 'x': equation.grid.solution_x, 'num_evals': num_evals})


def integrate_exact(equation, times=_DEFAULT_TIMES, warmup=0, integrate_method='RK23',
                    filter_interval=None):
  """integrate.py:282-293."""
  equation = equation.to_exact()
  return integrate(equation, exact_differentiator(equation), times, warmup,
                   integrate_method=integrate_method, filter_interval=filter_interval)


def integrate_baseline(equation, times=_DEFAULT_TIMES, warmup=0, accuracy_order=1,
                       integrate_method='RK23', exact_filter_interval=None):
  """integrate.py:296-308."""
  return integrate(equation, PolynomialDifferentiator(equation, accuracy_order), times, warmup,
                   integrate_method=integrate_method, filter_interval=exact_filter_interval)


def integrate_weno(equation, times=_DEFAULT_TIMES, warmup=0, integrate_method='RK23',
                   exact_filter_interval=None, **kwargs):
  """integrate.py:311-324."""
  if type(equation) not in equations_lib.FLUX_EQUATION_TYPES.values():
    raise ValueError('invalid equation: {}'.format(equation))
  return integrate(equation, WENODifferentiator(equation, **kwargs), times, warmup,
                   integrate_method=integrate_method, filter_interval=exact_filter_interval)


def integrate_spectral(equation, times=_DEFAULT_TIMES, warmup=0, integrate_method='RK23',
                       exact_filter_interval=None):
  """integrate.py:327-339."""
  if type(equation) not in equations_lib.EQUATION_TYPES.values():
    raise ValueError('invalid equation: {}'.format(equation))
  return integrate(equation, SpectralDifferentiator(equation), times, warmup,
                   integrate_method=integrate_method, filter_interval=exact_filter_interval)


def integrate_exact_baseline_and_model(checkpoint_dir, hparams=None, random_seed=0,
                                       times=_DEFAULT_TIMES, warmup=0, integrate_method='RK23',
                                       exact_filter_interval=None):
  """Exact fine-grid run, then baseline and learned model on the coarse grid from the
  same resampled initial condition (integrate.py:342-396)."""
  if hparams is None:
    hparams = training.load_hparams(checkpoint_dir)        # integrate.py:349-350
  fine, coarse = equations_lib.from_hparams(hparams, random_seed=random_seed)
  exact = integrate_exact(fine, times, warmup, integrate_method=integrate_method,
                          filter_interval=exact_filter_interval)
  solution_exact = np.asarray(exact['y'].data)
  y0 = coarse.grid.resample(solution_exact[0, :])
  if np.isnan(y0).any():
    raise ValueError('solution contains NaNs')
  baseline, evals_baseline = odeint(y0, PolynomialDifferentiator(coarse), warmup + times,
                                    method=integrate_method)
  learned, evals_model = odeint(y0, SavedModelDifferentiator(checkpoint_dir, coarse, hparams),
                                warmup + times, method=integrate_method)
  return make_dataset(
      {'y_exact': (('time', 'x_high'), solution_exact),
       'y_baseline': (('time', 'x_low'), baseline),
       'y_model': (('time', 'x_low'), learned)},
      {'time': warmup + times, 'x_low': coarse.grid.solution_x, 'x_high': fine.grid.solution_x,
       'num_evals_exact': int(np.asarray(exact['num_evals'].data if hasattr(exact['num_evals'], 'data')
                                         else exact['num_evals'])),
       'num_evals_baseline': evals_baseline, 'num_evals_model': evals_model})


def integrate_model_from_warm_start(checkpoint_dir, y0, hparams=None, random_seed=0,
                                    times=_DEFAULT_TIMES, warmup=0, integrate_method='RK23'):
  """integrate.py:399-427."""
  if hparams is None:
    hparams = training.load_hparams(checkpoint_dir)        # integrate.py:406-407
  _, coarse = equations_lib.from_hparams(hparams, random_seed=random_seed)
  solution, num_evals = odeint(y0, SavedModelDifferentiator(checkpoint_dir, coarse, hparams),
                               warmup + times, method=integrate_method)
  return make_dataset({'y': (('time', 'x'), solution)},
                      {'time': warmup + times, 'x': coarse.grid.solution_x, 'num_evals': num_evals})


# ---------------------------------------------------------------------------------
# Batched, fused fixed-step integration (the B200 path)
# ---------------------------------------------------------------------------------
class BatchIntegrator(object):
  """A batch of independent samples (one Equation per sample: they differ by
  random_seed only) advanced together by ddd1d_integrate."""

  def __init__(self, solver):
    self.solver = solver
    self.equations = solver.equations
    self.equation = solver.equation

  @classmethod
  def learned(cls, equations, hparams, weights, **kwargs):
    return cls(runtime.learned_solver(equations, hparams, _load_weights(weights), **kwargs))

  @classmethod
  def baseline(cls, equations, accuracy_order=1, **kwargs):
    return cls(runtime.stencil_solver(equations, accuracy_order, **kwargs))

  @classmethod
  def weno(cls, equations, non_weno_accuracy_order=3, **kwargs):
    return cls(runtime.weno_solver(equations, non_weno_accuracy_order, **kwargs))

  def initial_values(self):
    return np.stack([e.initial_value() for e in self.equations]).astype(np.float32)

  def rhs(self, t, u, sample_offset=0):
    return self.solver.rhs(t, u, sample_offset)

  def integrate(self, u0, t0=0.0, dt=None, num_steps=1, save_every=1, scheme='rk3', sample_offset=0,
                return_first_bad=False):
    """Device tensor [num_steps // save_every, batch, x]."""
    dt = self.equation.time_step if dt is None else dt
    return self.solver.integrate(u0, t0, dt, num_steps, save_every, scheme, sample_offset,
                                 return_first_bad)

  def odeint(self, u0=None, times=_DEFAULT_TIMES, method='RK23', rtol=1e-3, atol=1e-6, max_step=0.01):
    """Batched twin of odeint(): every sample runs SciPy's adaptive RK23 (same controller,
    FSAL, dense output, NaN padding) inside one kernel launch.
    Returns (y float64 [sample, time, x], nfev [sample])."""
    if method != 'RK23':
      raise NotImplementedError('only RK23 (the reference default, integrate.py:146) runs on the device')
    u0 = np.stack([e.initial_value() for e in self.equations]) if u0 is None else u0
    y, nfev, _ = self.solver.odeint(u0, times, rtol, atol, max_step)
    return y.permute(1, 0, 2).cpu().numpy(), nfev.cpu().numpy()

  def integrate_adaptive(self, u0=None, times=_DEFAULT_TIMES, **kwargs):
    """{y: (sample, time, x)} + num_evals per sample, the batched form of integrate()."""
    y, nfev = self.odeint(u0, times, **kwargs)
    return make_dataset(
        {'y': (('sample', 'time', 'x'), y), 'num_evals': (('sample',), nfev)},
        {'time': np.asarray(times), 'x': self.equation.grid.solution_x,
         'sample': np.array([e.random_seed for e in self.equations][:y.shape[0]])})

  def integrate_times(self, u0=None, times=_DEFAULT_TIMES, dt=None, scheme='rk3'):
    """Fixed-step twin of integrate(): samples at `times` (uniformly spaced, spacing a
    multiple of dt).  Rows that diverged are NaN from the first bad sample on, like
    the reference's NaN padding (integrate.py:161-167).  Returns {y: (sample, time, x)}."""
    import torch
    times = np.asarray(times, dtype=np.float64)
    dt = self.equation.time_step if dt is None else dt
    spacing = np.diff(times)
    every = int(round(spacing[0] / dt))
    if every < 1 or abs(every * dt - spacing[0]) > 1e-9 * max(1.0, abs(spacing[0])) or \
        (abs(spacing - spacing[0]) > 1e-9).any():
      raise ValueError('times must be uniformly spaced by a multiple of dt={}'.format(dt))
    u0 = self.initial_values() if u0 is None else u0
    rows = self.solver._rows(u0)
    snaps, bad = self.solver.integrate(rows, times[0], dt, every * (len(times) - 1), every, scheme,
                                       return_first_bad=True)
    y = torch.cat([rows[None], snaps], dim=0)
    first_bad_save = torch.where(bad >= 0, bad // every + 1, torch.full_like(bad, len(times)))
    mask = torch.arange(len(times), device=y.device)[:, None] >= first_bad_save[None, :]
    y = torch.where(mask[:, :, None], torch.full_like(y, float('nan')), y)
    return make_dataset(
        {'y': (('sample', 'time', 'x'), y.permute(1, 0, 2).cpu().numpy())},
        {'time': times, 'x': self.equation.grid.solution_x,
         'sample': np.array([e.random_seed for e in self.equations][:rows.shape[0]])})
