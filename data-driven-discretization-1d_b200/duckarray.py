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


# ---- the remaining duck helpers of duckarray.py (NumPy arrays or torch tensors) ----
def get_shape(x):
  """duckarray.py:44-52."""
  return tuple(x.shape)


def stack(arrays, axis=0):
  if _is_torch(arrays[0]):
    import torch
    return torch.stack(list(arrays), dim=axis)
  return np.stack(arrays, axis=axis)


def reshape(x, shape):
  return x.reshape(shape)


def sin(x):
  if _is_torch(x):
    import torch
    return torch.sin(x)
  return np.sin(x)


def sum(x, axis=None, keepdims=False):   # pylint: disable=redefined-builtin
  if _is_torch(x):
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)
  return np.sum(x, axis=axis, keepdims=keepdims)


def mean(x, axis=None, keepdims=False):
  if _is_torch(x):
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)
  return np.mean(x, axis=axis, keepdims=keepdims)


def rfft(x):
  """duckarray.py:84-94 (last axis)."""
  if _is_torch(x):
    import torch
    return torch.fft.rfft(x)
  return np.fft.rfft(x)


def irfft(x, n=None):
  if _is_torch(x):
    import torch
    return torch.fft.irfft(x, n=n)
  return np.fft.irfft(x, n=n)


def spectral_derivative(x, order=1, period=2 * np.pi):
  """duckarray.py:105-112: derivative of a periodic signal through the FFT (cuFFT for CUDA tensors)."""
  n = x.shape[-1]
  if n % 2:
    raise ValueError('spectral derivative only works for even length data')
  if _is_torch(x):
    from . import model
    return model.spectral_derivative(x, order, period)
  c = 2 * np.pi * 1j / period
  k = np.fft.rfftfreq(n, d=1 / n)
  return np.fft.irfft((c * k) ** order * np.fft.rfft(x), n=n)


def smoothing_filter(x, alpha=-np.log(1e-15), order=2):
  """duckarray.py:115-128: exponential filter exp(-alpha (k / k_max)^(2 order)) in Fourier space."""
  n = x.shape[-1]
  if n % 2:
    raise ValueError('smoothing filter only works for even length data')
  if _is_torch(x):
    import torch
    k = torch.fft.rfftfreq(n, d=1.0 / n, device=x.device)
    eta = k / k.max()
    sigma = torch.exp(-alpha * eta ** (2 * order))
    return torch.fft.irfft(sigma * torch.fft.rfft(x), n=n)
  k = np.fft.rfftfreq(n, d=1 / n)
  eta = k / k.max()
  sigma = np.exp(-alpha * eta ** (2 * order))
  return np.fft.irfft(sigma * np.fft.rfft(x), n=n)
