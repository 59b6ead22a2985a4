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


# ---------------------------------------------------------------------------------
# The intermediate quantities of weno.py as tensor expressions (the fused kernels compute them in registers
# and never materialise them; these are for inspection and for callers of the reference's helper names).
# NumPy in -> NumPy out, torch in -> torch out (on the tensor's device).
# ---------------------------------------------------------------------------------
def _as_torch(u):
  import torch
  if isinstance(u, torch.Tensor):
    return u, False
  return torch.as_tensor(np.asarray(u)), True


def _back(x, was_numpy):
  return x.cpu().numpy() if was_numpy else x


def _indicators(u):
  import torch
  m2, m1, p1, p2 = (torch.roll(u, s, dims=-1) for s in (2, 1, -1, -2))
  return torch.stack([
      1 / 4 * (m2 - 4 * m1 + 3 * u) ** 2 + 13 / 12 * (m2 - 2 * m1 + u) ** 2,
      1 / 4 * (m1 - p1) ** 2 + 13 / 12 * (m1 - 2 * u + p1) ** 2,
      1 / 4 * (3 * u - 4 * p1 + p2) ** 2 + 13 / 12 * (u - 2 * p1 + p2) ** 2,
  ], dim=-2)


def _omega(u, optimal_linear_weights, epsilon, p):
  import torch
  d = torch.as_tensor(np.asarray(optimal_linear_weights, dtype=np.float64), device=u.device).to(u.dtype)
  alpha = d[:, None] / (epsilon + _indicators(u)) ** p
  return alpha / alpha.sum(dim=-2, keepdim=True)


def calculate_smoothness_indicators(u):
  """weno.py:43-57: [..., x] -> [..., 3, x]."""
  x, was_numpy = _as_torch(u)
  return _back(_indicators(x), was_numpy)


def calculate_omega(u, optimal_linear_weights=OPTIMAL_SMOOTH_WEIGHTS, epsilon=1e-6, p=2):
  """weno.py:60-73: nonlinear weights of the three sub-stencils, [..., 3, x]."""
  x, was_numpy = _as_torch(u)
  return _back(_omega(x, optimal_linear_weights, epsilon, p), was_numpy)


def left_coefficients(u):
  """weno.py:76-89: the five linear coefficients of the left-biased reconstruction, [..., x, 5]."""
  import torch
  x, was_numpy = _as_torch(u)
  w = _omega(x, OPTIMAL_SMOOTH_WEIGHTS, 1e-6, 2)
  w0, w1, w2 = w[..., 0, :], w[..., 1, :], w[..., 2, :]
  out = torch.stack([w0 / 3, -(7 * w0 + w1) / 6, (11 * w0 + 5 * w1 + 2 * w2) / 6, (2 * w1 + 5 * w2) / 6, -w2 / 6],
                    dim=-1)
  return _back(out, was_numpy)


def right_coefficients(u):
  """weno.py:100-115 (reversed optimal weights, omega rolled by -1), [..., x, 5]."""
  import torch
  x, was_numpy = _as_torch(u)
  w = torch.roll(_omega(x, OPTIMAL_SMOOTH_WEIGHTS[::-1], 1e-6, 2), -1, dims=-1)
  w2, w1, w0 = w[..., 0, :], w[..., 1, :], w[..., 2, :]
  out = torch.stack([-w2 / 6, (5 * w2 + 2 * w1) / 6, (2 * w2 + 5 * w1 + 11 * w0) / 6, -(w1 + 7 * w0) / 6, w0 / 3],
                    dim=-1)
  return _back(out, was_numpy)
