"""Array helpers shared by host-side code (NumPy arrays or torch tensors) --
the subset of pde_superresolution/duckarray.py that the integration path uses."""
import numpy as np


def _is_torch(x):
  return type(x).__module__.startswith('torch')


def _axis(axis, ndim):
  if not -ndim <= axis < ndim:
    raise ValueError('invalid axis {} for ndim {}'.format(axis, ndim))
  return axis % ndim


def resample_mean(inputs, factor, axis=-1):
  """Block means of `factor` consecutive points (duckarray.py:139-163)."""
  shape = tuple(inputs.shape)
  axis = _axis(axis, len(shape))
  if shape[axis] % factor:
    raise ValueError('resample factor {} must divide size {}'.format(factor, shape[axis]))
  blocked = shape[:axis] + (shape[axis] // factor, factor) + shape[axis + 1:]
  return inputs.reshape(blocked).mean(axis + 1)


def subsample(inputs, factor, axis=-1):
  """Every `factor`-th point (duckarray.py:166-189)."""
  shape = tuple(inputs.shape)
  axis = _axis(axis, len(shape))
  if shape[axis] % factor:
    raise ValueError('resample factor {} must divide size {}'.format(factor, shape[axis]))
  index = [slice(None)] * len(shape)
  index[axis] = slice(None, None, factor)
  return inputs[tuple(index)]


RESAMPLE_FUNCS = {'mean': resample_mean, 'subsample': subsample}


def roll(tensor, shift, axis):
  """Periodic shift (duckarray.py:206-219)."""
  if _is_torch(tensor):
    import torch
    return torch.roll(tensor, shift, axis)
  return np.roll(tensor, shift, axis)


def where(cond, x, y):
  if _is_torch(cond):
    import torch
    return torch.where(cond, x, y)
  return np.where(cond, x, y)


def minimum(x, y):
  if _is_torch(x):
    import torch
    return torch.minimum(x, y)
  return np.minimum(x, y)


def maximum(x, y):
  if _is_torch(x):
    import torch
    return torch.maximum(x, y)
  return np.maximum(x, y)


def concatenate(arrays, axis):
  if _is_torch(arrays[0]):
    import torch
    return torch.cat(list(arrays), dim=axis)
  return np.concatenate(arrays, axis=axis)
