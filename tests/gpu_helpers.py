"""Helpers for the -m gpu tests: build product-side objects next to oracle-side ones."""
import json

import numpy as np

from oracle import pde_oracle as O

REG = None


def registry():
  global REG
  if REG is None:
    from ddd1d_b200 import equations
    REG = {'plain': equations.EQUATION_TYPES, 'conservative': equations.CONSERVATIVE_EQUATION_TYPES,
           'godunov': equations.FLUX_EQUATION_TYPES}
  return REG


def product_equation(kind, variant, n, seed=0, resample_factor=1):
  return registry()[variant][kind](n, resample_factor=resample_factor, random_seed=seed)


def product_hparams(kind, variant, n, resample_factor=1, **overrides):
  from ddd1d_b200 import training
  return training.create_hparams(
      kind, conservative=variant != 'plain', numerical_flux=variant == 'godunov',
      resample_factor=resample_factor,
      equation_kwargs=json.dumps({'num_points': n * resample_factor}), **overrides)


def oracle_equation(kind, variant, n, seed=0, resample_factor=1):
  return O.EquationSpec(kind, variant, num_points=n, random_seed=seed, resample_factor=resample_factor)


def smooth_rows(batch, n, seed=0, amplitude=0.6):
  """Smooth periodic rows (a few Fourier modes) -- O(1) fields like real solutions."""
  rs = np.random.RandomState(seed)
  x = 2 * np.pi * np.arange(n) / n
  out = np.zeros((batch, n))
  for m in range(1, 5):
    out += rs.randn(batch, 1) * np.sin(m * x + 2 * np.pi * rs.rand(batch, 1)) / m
  return (amplitude * out).astype(np.float32)
