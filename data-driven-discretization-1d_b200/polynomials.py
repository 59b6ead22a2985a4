"""Finite-difference / finite-volume stencil tables and the polynomial-accuracy
projection -- host-side mirror of pde_superresolution/polynomials.py.

Table construction is one-time NumPy work on the host, exactly as in the
reference (it runs at graph-build time there, polynomials.py:209-264); applying
the tables to data happens inside the CUDA row kernel.

Compatibility note: trained checkpoints are only meaningful relative to the
null-space basis the reference computed with ``np.linalg.svd`` of the constraint
matrix (polynomials.py:246-254).  ``constraints`` therefore reproduces the
reference's matrix bit for bit (same rows, same order, same arithmetic) so the
same LAPACK call returns the same basis; tests/test_host_tables.py checks this
against fixtures minted from the reference.
"""
import enum
import math

import numpy as np

from . import _lib


class GridOffset(enum.Enum):
  """Where stencil outputs sit relative to the input grid (polynomials.py:31-34)."""
  CENTERED = 1
  STAGGERED = 2


class Method(enum.Enum):
  """polynomials.py:37-40."""
  FINITE_DIFFERENCES = 1
  FINITE_VOLUMES = 2


def regular_grid(grid_offset, derivative_order, accuracy_order=1, dx=1):
  """Smallest regular stencil supporting the requested orders (polynomials.py:43-71)."""
  need = derivative_order + accuracy_order
  if grid_offset is GridOffset.CENTERED:
    reach = need // 2
    return dx * np.arange(-reach, reach + 1)
  if grid_offset is GridOffset.STAGGERED:
    reach = (need + 1) // 2
    return dx * (np.arange(-reach, reach) + 0.5)
  raise ValueError('unexpected grid_offset: {}'.format(grid_offset))


def _moment_row(grid, delta, method, power):
  if method is Method.FINITE_DIFFERENCES:
    return grid ** power
  if method is Method.FINITE_VOLUMES:
    # cell average of x**power over [x - delta/2, x + delta/2]
    return (1 / delta * ((grid + delta / 2) ** (power + 1) - (grid - delta / 2) ** (power + 1))
            / (power + 1))
  raise ValueError('unexpected method: {}'.format(method))


def constraints(grid, method, derivative_order, accuracy_order=None):
  """Linear system A c = b characterising stencils of the given accuracy
  (polynomials.py:74-149).  Rows: the vanishing moments, de-duplicated and sorted,
  then the moment that must equal derivative_order!."""
  grid = np.asarray(grid)
  if accuracy_order is None:
    accuracy_order = grid.size - derivative_order
  if accuracy_order < 1:
    raise ValueError('cannot compute constriants with non-positive accuracy_order: {}'
                     .format(accuracy_order))
  spacings = np.unique(np.diff(grid))
  if (abs(spacings - spacings[0]) > 1e-8).any():
    raise ValueError('not a regular grid: {}'.format(spacings))
  delta = spacings[0]
  vanishing, pinned = set(), None
  for power in range(accuracy_order + derivative_order):
    row = _moment_row(grid, delta, method, power)
    if power == derivative_order:
      pinned = row
    else:
      vanishing.add(tuple(row))
  if len(vanishing) + 1 > grid.size:
    raise ValueError('no valid {} stencil exists for derivative_order={} and accuracy_order={} '
                     'with grid={}'.format(method, derivative_order, accuracy_order, grid))
  matrix = np.array(sorted(vanishing) + [pinned])
  rhs = np.zeros(matrix.shape[0])
  rhs[-1] = math.factorial(derivative_order)
  return matrix, rhs


def coefficients(grid, method, derivative_order):
  """The unique maximal-accuracy stencil on `grid` (polynomials.py:152-167)."""
  matrix, rhs = constraints(grid, method, derivative_order)
  return np.linalg.solve(matrix, rhs)


def zero_padded_coefficients(grid, method, derivative_order, padding):
  """Standard coefficients on the grid trimmed by `padding`, zero-extended back
  (polynomials.py:170-195)."""
  lead, trail = padding
  inner = np.asarray(grid)[lead:len(grid) - trail]
  return np.pad(coefficients(inner, method, derivative_order), (lead, trail), mode='constant')


class PolynomialAccuracyLayer(object):
  """Affine map z -> bias + z @ nullspace whose image satisfies the accuracy
  constraints (polynomials.py:198-277).

  Attributes: input_size, grid_size, bias [grid_size], nullspace [input_size, grid_size].
  """

  def __init__(self, grid, method, derivative_order, accuracy_order=2, bias=None,
               bias_zero_padding=(0, 0), out_scale=1.0):
    grid = np.asarray(grid)
    matrix, rhs = constraints(grid, method, derivative_order, accuracy_order)
    if bias is None:
      bias = zero_padded_coefficients(grid, method, derivative_order, bias_zero_padding)
    if np.linalg.norm(matrix.dot(bias) - rhs) > 1e-8:
      raise ValueError('invalid bias, not in nullspace')
    free = matrix.shape[1] - matrix.shape[0]
    if not free:
      raise ValueError('there is only one valid solution accurate to this order')
    basis = np.linalg.svd(matrix)[2][-free:]
    spacing = grid[1] - grid[0]
    self.input_size = free
    self.grid_size = grid.size
    self.nullspace = basis * (out_scale / spacing ** derivative_order)
    self.bias = bias

  def apply(self, inputs):
    """[batch, x, input_size] -> [batch, x, grid_size] (float32 tables as in
    polynomials.py:275-277).  Torch tensors in, torch tensors out; inside the
    integrator this projection is fused into the row kernel instead."""
    import torch
    bias = torch.as_tensor(self.bias.astype(np.float32), device=inputs.device)
    basis = torch.as_tensor(self.nullspace.astype(np.float32), device=inputs.device)
    return bias + torch.einsum('bxi,ij->bxj', inputs, basis)

  # window forms consumed by libddd1d (include/ddd1d.h: "Stencil window")
  def bias_window(self):
    return _lib.to_window(self.bias)

  def nullspace_window(self):
    return _lib.to_window(self.nullspace)


def reconstruct(inputs, grid, method, derivative_order):
  """Apply the standard stencil for `grid` along x with periodic wrap
  (polynomials.py:280-303).  inputs: [batch, x] CUDA tensor or array; runs on the GPU."""
  from . import model
  return model.apply_fixed_stencils(inputs, [coefficients(grid, method, derivative_order)])[..., 0]
