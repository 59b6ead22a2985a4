"""Evaluation of a trained model over many samples -- the callers and data formats on
the far side of the hot path:

* `run_integrate_batch`: scripts/run_evaluation.py:152-174 (`run_integrate`, one Beam
  `Map` element per seed there) for ALL seeds in one launch of the device-side adaptive
  RK23 (or the fused fixed-step integrator);
* `unify_x_coords`, `calculate_mae`, `is_good`, `mostly_good`, `calculate_survival`,
  `mostly_good_survival`: analysis.py:39-90 and scripts/run_evaluation.py:189-210 as
  torch reductions over the gathered trajectories (they run where the tensors live);
* `write_results` / `read_results`: `results.nc` as the reference writes it -- `Dataset.to_netcdf()` bytes
  (xarray_beam.py:32-35), i.e. NetCDF-3 through xarray's SciPy backend -- with the schema
  `{y: (sample, time, x)}` + coords `time, x, sample`, `num_evals(sample)` (scripts/run_evaluation.py:168-174,
  186-187), written with `scipy.io.netcdf_file`; `write_metric` does the same for `mae.nc` / `survival.nc`
  (scripts/run_evaluation.py:189-210).  A path ending in `.npz` keeps the NumPy container;
* `run_integrate_batch(..., distributed=True)`: the per-seed map of scripts/run_evaluation.py:212-221 over
  the GPUs of one box -- every rank integrates a contiguous block of seeds, one all-gather at the end.

Arrays may be NumPy or torch (CPU or CUDA); results come back as NumPy.
"""
import json

import numpy as np


def _t(x, like=None):
  import torch
  if isinstance(x, torch.Tensor):
    return x
  return torch.as_tensor(np.asarray(x), device=None if like is None else like.device)


def run_integrate_batch(checkpoint_dir, hparams, initial_conditions, times, warmup=0.0,
                        integrate_method='RK23', fixed_dt=None, first_seed=0, distributed=False, group=None):
  """Integrate the learned model from `initial_conditions[sample, x]`, sample i with
  `random_seed = first_seed + i` (scripts/run_evaluation.py:147-166).

  integrate_method='RK23' runs SciPy's adaptive scheme per row on the device (same
  controller and dense output as integrate.odeint); `fixed_dt` switches to the fused
  fixed-step Bogacki-Shampine integrator.  Returns the results dict of write_results.

  distributed=True (inside an initialised torch.distributed job, one rank per GPU): every rank passes the
  SAME full `initial_conditions`, integrates its contiguous block of samples (distributed.shard_bounds) and
  receives the results of all samples -- the reference's Beam map over seeds (scripts/run_evaluation.py:
  212-221) with one all-gather at the end.  Rows are independent, so the result is bit-identical to the
  single-GPU call."""
  from . import equations as equations_lib
  from . import integrate
  y0 = np.asarray(initial_conditions)
  if y0.ndim != 2:
    raise ValueError('initial_conditions must be [sample, x]')
  if np.isnan(y0).any():
    raise ValueError('initial conditions cannot have NaNs')        # run_evaluation.py:142-143
  if hparams is None:
    from . import training
    hparams = training.load_hparams(checkpoint_dir)
  if distributed:
    return _run_sharded(checkpoint_dir, hparams, y0, times, warmup, integrate_method, fixed_dt, first_seed, group)
  coarse = [equations_lib.from_hparams(hparams, random_seed=first_seed + i)[1] for i in range(y0.shape[0])]
  weights = integrate._load_weights(checkpoint_dir)
  batch = integrate.BatchIntegrator.learned(coarse, hparams, weights)
  times = np.asarray(times, dtype=np.float64)
  if fixed_dt is None:
    y, nfev = batch.odeint(y0, warmup + times, method=integrate_method)
  else:
    ds = batch.integrate_times(y0, warmup + times, dt=fixed_dt)
    y = np.asarray(ds['y'].data)
    steps = int(round((times[-1] - times[0]) / fixed_dt))
    nfev = np.full(y0.shape[0], 3 * steps, dtype=np.int64)
  return {'y': np.asarray(y), 'time': warmup + times, 'x': coarse[0].grid.solution_x,
          'num_evals': np.asarray(nfev), 'sample': first_seed + np.arange(y0.shape[0])}


def _run_sharded(checkpoint_dir, hparams, y0, times, warmup, integrate_method, fixed_dt, first_seed, group,
                 local_runner=None):
  """The distributed branch of run_integrate_batch.  `local_runner` (tests: the CPU plumbing under gloo)
  replaces the CUDA integration of one block."""
  import torch
  import torch.distributed as dist
  from . import distributed as D
  if not (dist.is_available() and dist.is_initialized()):
    raise RuntimeError('distributed=True needs an initialised torch.distributed process group')
  rank, world = dist.get_rank(group), dist.get_world_size(group)
  total = y0.shape[0]
  start, stop = D.shard_bounds(total, rank, world)
  run = local_runner or (lambda block, seed0: run_integrate_batch(
      checkpoint_dir, hparams, block, times, warmup, integrate_method, fixed_dt, seed0))
  times = np.asarray(times, dtype=np.float64)
  if stop > start:
    local = run(y0[start:stop], first_seed + start)
    y_local, evals_local = np.asarray(local['y']), np.asarray(local['num_evals'])
    x = np.asarray(local['x'])
  else:                                               # more ranks than samples: an empty block
    y_local = np.zeros((0, len(times), y0.shape[1]))
    evals_local = np.zeros((0,), dtype=np.int64)
    x = None
  device = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
  y = D.gather_snapshots(torch.as_tensor(y_local, dtype=torch.float64, device=device), total, sample_axis=0, group=group)
  evals = D.gather_snapshots(torch.as_tensor(evals_local.astype(np.int64), device=device)[:, None], total,
                             sample_axis=0, group=group)[:, 0]
  if x is None:                                       # the grid is a function of hparams alone
    from . import equations as equations_lib
    x = equations_lib.from_hparams(hparams, random_seed=first_seed)[1].grid.solution_x
  return {'y': y.cpu().numpy(), 'time': warmup + times, 'x': x, 'num_evals': evals.cpu().numpy(),
          'sample': first_seed + np.arange(total)}


def _to_netcdf3(path, dims, variables, attrs=None):
  """Write NetCDF-3 (classic, 64-bit offsets) with SciPy: `dims` {name: size}, `variables`
  {name: (dim names, array)}.  What xarray's Dataset.to_netcdf() emits with its SciPy backend, the
  backend the reference's bytes-in-memory writer uses (xarray_beam.py:32-35)."""
  from scipy.io import netcdf_file
  with netcdf_file(path, 'w', version=2) as f:
    for name, size in dims.items():
      f.createDimension(name, int(size))
    for name, spec in variables.items():
      vdims, array = spec[0], np.asarray(spec[1])
      if array.dtype == np.int64:                     # NetCDF-3 has no 64-bit integers (xarray casts likewise)
        array = array.astype(np.int32)
      elif array.dtype == np.bool_:
        array = array.astype(np.int8)
      var = f.createVariable(name, array.dtype, tuple(vdims))
      var[...] = array
      for key, value in (spec[2] if len(spec) > 2 else {}).items():
        setattr(var, key, value)
    for key, value in (attrs or {}).items():
      setattr(f, key, value)


def write_results(path, results):
  """`results.nc`: y (sample, time, x) + coords time, x, sample and num_evals (sample), NetCDF-3; a path
  ending in .npz keeps the NumPy container."""
  y = np.asarray(results['y'])
  if y.ndim != 3 or y.shape != (len(results['sample']), len(results['time']), len(results['x'])):
    raise ValueError('y must be (sample, time, x)')
  if str(path).endswith('.npz'):
    np.savez_compressed(path, y=y, time=np.asarray(results['time']), x=np.asarray(results['x']),
                        num_evals=np.asarray(results['num_evals']), sample=np.asarray(results['sample']),
                        dims=json.dumps({'y': ['sample', 'time', 'x'], 'num_evals': ['sample']}))
    return
  _to_netcdf3(path, {'sample': y.shape[0], 'time': y.shape[1], 'x': y.shape[2]},
              {'y': (('sample', 'time', 'x'), y, {'coordinates': 'num_evals'}),     # xarray's encoding of a non-index coordinate
               'time': (('time',), np.asarray(results['time'], dtype=np.float64)),
               'x': (('x',), np.asarray(results['x'], dtype=np.float64)),
               'sample': (('sample',), np.asarray(results['sample'])),
               'num_evals': (('sample',), np.asarray(results['num_evals']))})


def read_results(path):
  if str(path).endswith('.npz'):
    with np.load(path) as f:
      return {k: f[k] for k in ('y', 'time', 'x', 'num_evals', 'sample')}
  from scipy.io import netcdf_file
  with netcdf_file(path, 'r', mmap=False) as f:
    return {k: np.array(f.variables[k][...]) for k in ('y', 'time', 'x', 'num_evals', 'sample')}


def write_metric(path, name, values, stop_times=None, samples=None):
  """`mae.nc` (name='mae': values [time_max, sample], scripts/run_evaluation.py:189-198) or `survival.nc`
  (name='survival': values [sample], :200-210) as NetCDF-3."""
  values = np.asarray(values, dtype=np.float64)
  if values.ndim == 2:
    stop_times = np.arange(values.shape[0]) if stop_times is None else stop_times
    samples = np.arange(values.shape[1]) if samples is None else samples
    _to_netcdf3(path, {'time_max': values.shape[0], 'sample': values.shape[1]},
                {name: (('time_max', 'sample'), values), 'time_max': (('time_max',), np.asarray(stop_times, np.float64)),
                 'sample': (('sample',), np.asarray(samples))})
  elif values.ndim == 1:
    samples = np.arange(values.shape[0]) if samples is None else samples
    _to_netcdf3(path, {'sample': values.shape[0]},
                {name: (('sample',), values), 'sample': (('sample',), np.asarray(samples))})
  else:
    raise ValueError('metric must be [time_max, sample] or [sample]')


# -------------------------------------------------------------------------------------
# analysis.py
# -------------------------------------------------------------------------------------
def unify_x_coords(y_high, factor):
  """analysis.py:39-52: bring a fine-grid variable to the coarse grid by block means
  (duckarray.resample_mean, duckarray.py:139-163) along the last axis."""
  y = _t(y_high)
  if y.shape[-1] % factor:
    raise ValueError('resample factor {} must divide size {}'.format(factor, y.shape[-1]))
  return y.reshape(y.shape[:-1] + (y.shape[-1] // factor, factor)).mean(dim=-1)


def calculate_mae(y_model, y_exact, times, stop_times):
  """scripts/run_evaluation.py:189-198: mean |model - exact| over x and over the times up
  to each `time_max` (label-inclusive slice), NaNs propagating (skipna=False).
  y_*: (sample, time, x) on the same grid.  Returns float64 [time_max, sample]."""
  import torch
  model, exact = _t(y_model), _t(y_exact)
  exact = exact.to(model.device)
  err = (model - exact).abs().to(torch.float64).mean(dim=-1)             # (sample, time)
  times = np.asarray(times, dtype=np.float64)
  out = []
  for time_max in stop_times:
    count = int(np.searchsorted(times, time_max, side='right'))
    if count == 0:
      out.append(torch.full((err.shape[0],), float('nan'), dtype=torch.float64, device=err.device))
    else:
      out.append(err[:, :count].mean(dim=1))
  return torch.stack(out).cpu().numpy()


def is_good(model, exact, max_error=0.5):
  """analysis.py:55-61."""
  return (_t(model) - _t(exact).to(_t(model).device)).abs() <= max_error


def mostly_good(model, exact, max_error=0.5, frac_good=0.8):
  """analysis.py:64-71: per (sample, time), is the fraction of accurate points >= frac_good."""
  import torch
  return is_good(model, exact, max_error).to(torch.float64).mean(dim=-1) >= frac_good


def calculate_survival(good, times):
  """analysis.py:74-78: the "lifetime" of boolean series along the last (time) axis:
  times.max() if always true, else the time of the first False."""
  import torch
  good = _t(good).to(torch.bool)
  t = torch.as_tensor(np.asarray(times, dtype=np.float64), device=good.device)
  first_false = torch.argmin(good.to(torch.int8), dim=-1)               # first minimum = first False
  return torch.where(good.all(dim=-1), t.max().expand(first_false.shape), t[first_false]).cpu().numpy()


def mostly_good_survival(y_model, y_exact_high, times, quantile=0.8):
  """analysis.py:81-90 for one model variable: the error threshold is the (1 - quantile)
  quantile of |y_exact| on ITS OWN (fine) grid; survival is judged on the coarse grid.
  y_model (sample, time, x_low), y_exact_high (sample, time, x_high).  Returns [sample]."""
  import torch
  model = _t(y_model)
  exact = _t(y_exact_high).to(model.device)
  flat = exact.abs().to(torch.float64).flatten()
  # DataArray.quantile: np.nanpercentile semantics -- NaNs are skipped, linear interpolation between order
  # statistics
  flat = flat[~torch.isnan(flat)]
  if flat.numel() == 0:
    max_error = float('nan')
  else:
    srt = torch.sort(flat).values
    pos = (1.0 - quantile) * (srt.numel() - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, srt.numel() - 1)
    max_error = float(srt[lo] + (srt[hi] - srt[lo]) * (pos - lo))
  factor = exact.shape[-1] // model.shape[-1]
  good = mostly_good(model, unify_x_coords(exact, factor), max_error=max_error, frac_good=quantile)
  return calculate_survival(good, times)
