"""Periodic padding / fixed-filter convolution -- mirror of the stand-alone helpers of
pde_superresolution/layers.py.

Inside the integrator none of this exists as a separate op: the periodic halo lives in
shared memory (or in the tensor engine's activation planes) and the conv stack is fused
(csrc/).  These functions serve callers that use the helpers on their own
(polynomials.reconstruct, alignment checks).  `conv1d_periodic_layer` needs trainable
variables in the reference (tf.layers.conv1d); here the conv stack takes its weights
explicitly through model.predict_coefficients, so no stand-alone layer is offered.
"""
import numpy as np


def _torch(x):
  import torch
  return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))


def static_or_dynamic_size(tensor, axis):
  """layers.py:26-36."""
  shape = tuple(tensor.shape)
  if not -len(shape) <= axis < len(shape):
    raise ValueError('axis {} out of bounds for tensor of rank {}'.format(axis, len(shape)))
  return shape[axis]


def pad_periodic(inputs, padding, center=False, name=None):
  """[batch, length, features] -> [batch, length + padding, features] with periodic wrap
  (layers.py:39-83): center=False appends `padding` points on the right; center=True puts
  ceil(padding / 2) on the left (`-padding // 2`) and floor(padding / 2) on the right, tiling
  as often as needed."""
  del name
  x = _torch(inputs)
  if x.dim() != 3:
    raise ValueError('inputs must be 3D for periodic padding')
  if padding == 0:
    return x
  import torch
  n = x.shape[1]
  if center:
    left = -(-padding // 2)
    index = torch.arange(-left, n + padding // 2, device=x.device) % n
  else:
    index = torch.arange(0, n + padding, device=x.device) % n
  return x.index_select(1, index)


def nn_conv1d_periodic(inputs, filters, stride=1, center=False):
  """tf.nn.conv1d with periodic boundary conditions (layers.py:95-100) for the case the
  reference uses it for: one input and one output channel (polynomials.reconstruct).
  inputs [batch, length, 1], filters [width, 1, 1].  center=True runs the CUDA library's
  fixed-stencil kernel (width <= 7); center=False is the same cross-correlation shifted."""
  import torch
  x = _torch(inputs)
  f = np.asarray(filters.detach().cpu() if isinstance(filters, torch.Tensor) else filters, dtype=np.float64)
  if x.dim() != 3 or x.shape[2] != 1 or f.ndim != 3 or f.shape[1:] != (1, 1):
    raise NotImplementedError('nn_conv1d_periodic is built for single-channel inputs and filters')
  if stride != 1:
    raise NotImplementedError('stride != 1')
  from . import model
  width = f.shape[0]
  out = model.apply_fixed_stencils(x[..., 0], [f[:, 0, 0]])            # centred: ceil((w-1)/2) points on the left
  if not center:
    out = torch.roll(out, -(-(-(width - 1) // 2)), dims=1)             # window starts AT the output point
  return out
