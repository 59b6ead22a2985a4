"""WENO5 reconstruction -- mirror of pde_superresolution/weno.py, evaluated by the
CUDA library (float32 or float64 following the input dtype)."""
import ctypes

import numpy as np

from . import _lib

OPTIMAL_SMOOTH_WEIGHTS = (0.1, 0.6, 0.3)


def _reconstruct(u):
  import torch
  if not torch.cuda.is_available():
    raise RuntimeError('ddd1d_b200 needs a CUDA device; there is no CPU fallback')
  was_numpy = not isinstance(u, torch.Tensor)
  x = torch.as_tensor(u)
  if x.dtype not in (torch.float32, torch.float64):
    x = x.to(torch.float64)
  shape = x.shape
  rows = x.reshape(-1, shape[-1]).cuda().contiguous()
  left, right = torch.empty_like(rows), torch.empty_like(rows)
  lib = _lib.load()
  real = _lib.REAL_F64 if rows.dtype == torch.float64 else _lib.REAL_F32
  stream = ctypes.c_void_p(torch.cuda.current_stream(rows.device).cuda_stream)
  _lib.check(lib.ddd1d_weno_reconstruct(rows.device.index, real, rows.data_ptr(), left.data_ptr(),
                                        right.data_ptr(), rows.shape[0], rows.shape[1], stream))
  left, right = left.reshape(shape), right.reshape(shape)
  if was_numpy:
    return left.cpu().numpy(), right.cpu().numpy()
  return left, right


def reconstruct_left(u):
  """u at x + 1/2 from the left-biased stencil, [..., x] (weno.py:92-97)."""
  return _reconstruct(u)[0]


def reconstruct_right(u):
  """u at x + 1/2 from the right-biased stencil, [..., x] (weno.py:118-123)."""
  return _reconstruct(u)[1]


def reconstruct_both(u):
  """(left, right) from one kernel launch."""
  return _reconstruct(u)
