"""Equation definitions -- host-side mirror of pde_superresolution/equations.py.

Same public names and call signatures as the reference (Grid, Equation and its
nine concrete classes, RandomForcing, the three registries, from_hparams) so code
written against the reference keeps working; the arithmetic that matters for
integration (equation_of_motion, forcing, flux differences) is evaluated inside
the CUDA row kernel, selected by the (KIND, VARIANT) tags each class carries.
The NumPy / torch implementations here serve initial conditions, host-side
checks and API parity.
"""
import enum
import json

import numpy as np

from . import duckarray
from . import polynomials


@enum.unique
class ExactMethod(enum.Enum):
  """How the reference produces "exact" fine-grid solutions (equations.py:37-41)."""
  POLYNOMIAL = 1
  SPECTRAL = 2
  WENO = 3


class Grid(object):
  """Solution grid, the finer reference grid, and the map between them
  (equations.py:44-68)."""

  def __init__(self, solution_num_points, resample_factor=1, resample_method='mean', period=1.0):
    self.resample_factor = resample_factor
    self.resample_method = resample_method
    self.period = period
    self.solution_num_points = solution_num_points
    self.solution_dx = period / solution_num_points
    self.solution_x = self.solution_dx * np.arange(solution_num_points)
    self.reference_num_points = solution_num_points * resample_factor
    self.reference_dx = period / self.reference_num_points
    self.reference_x = self.reference_dx * np.arange(self.reference_num_points)

  def resample(self, x, axis=-1):
    return duckarray.RESAMPLE_FUNCS[self.resample_method](x, self.resample_factor, axis=axis)


class RandomForcing(object):
  """Sum of `nparams` travelling sinusoids with seeded random parameters
  (equations.py:196-227).  The draw order (a, omega, k, phi from one RandomState)
  is part of the contract: seed s must give the reference's sample s."""

  def __init__(self, grid, nparams=20, seed=0, amplitude=1, k_min=1, k_max=3):
    self.grid = grid
    rng = np.random.RandomState(seed)
    self.a = 0.5 * amplitude * rng.uniform(-1, 1, size=(nparams, 1))
    self.omega = rng.uniform(-0.4, 0.4, size=(nparams, 1))
    wavenumbers = np.arange(k_min, k_max + 1)
    self.k = rng.choice(np.concatenate([-wavenumbers, wavenumbers]), size=(nparams, 1))
    self.phi = rng.uniform(0, 2 * np.pi, size=(nparams, 1))

  def __call__(self, t):
    phase = self.omega * t + 2 * np.pi * self.k * self.grid.reference_x / self.grid.period + self.phi
    return self.grid.resample(np.sum(self.a * np.sin(phase), axis=0))

  def export(self, path):
    header = np.zeros_like(self.a)
    header[0] = self.grid.period
    header[1] = self.grid.reference_num_points
    np.savetxt(path, np.array([self.a, self.omega, self.k, self.phi, header]).squeeze())


def staggered_first_derivative(y, dx):
  """(y[x+1] - y[x]) / dx with periodic wrap (equations.py:305-320)."""
  forward = duckarray.concatenate([y[..., 1:], y[..., :1]], axis=-1)
  return (1 / dx) * (forward - y)


def godunov_convective_flux(u_minus, u_plus):
  """Godunov flux of u**2/2 (equations.py:341-349)."""
  lo, hi = u_minus ** 2, u_plus ** 2
  return 0.5 * duckarray.where(u_minus <= u_plus, duckarray.minimum(lo, hi), duckarray.maximum(lo, hi))


class Equation(object):
  """Base class (equations.py:71-193).  Class attributes as in the reference plus
  KIND / VARIANT, the tags the CUDA library dispatches on."""
  CONSERVATIVE = ...
  GRID_OFFSET = ...
  EXACT_METHOD = ...
  DERIVATIVE_NAMES = ...
  DERIVATIVE_ORDERS = ...
  KIND = ...
  VARIANT = 'plain'
  FORCED = False          # finalize_time_derivative adds forcing(t)
  NUM_FORCING_TERMS = 20

  def __init__(self, num_points, resample_factor=1, period=1.0, random_seed=0, k_min=1, k_max=3):
    method = 'mean' if self.CONSERVATIVE else 'subsample'
    self.grid = Grid(num_points, resample_factor, method, period)
    self.random_seed = random_seed
    self.k_min = k_min
    self.k_max = k_max
    self.forcing = RandomForcing(self.grid, nparams=self.NUM_FORCING_TERMS, seed=random_seed,
                                 k_min=k_min, k_max=k_max)

  # -- to be provided by concrete equations ---------------------------------------
  def initial_value(self):
    raise NotImplementedError

  @property
  def time_step(self):
    raise NotImplementedError

  @property
  def standard_deviation(self):
    raise NotImplementedError

  def equation_of_motion(self, y, spatial_derivatives):
    raise NotImplementedError

  @classmethod
  def base_type(cls):
    raise NotImplementedError

  # -- shared behaviour ---------------------------------------------------------------
  def finalize_time_derivative(self, t, y_t):
    return y_t + self.forcing(t) if self.FORCED else y_t

  def params(self):
    return dict(num_points=self.grid.reference_num_points, period=self.grid.period,
                random_seed=self.random_seed, k_min=self.k_min, k_max=self.k_max)

  def to_fine(self):
    return type(self)(**self.params())

  @classmethod
  def exact_type(cls):
    return cls.base_type()

  @classmethod
  def conservative_type(cls):
    return CONSERVATIVE_EQUATION_TYPES[cls.KIND]

  def to_exact(self):
    return self.exact_type()(**self.params())

  def to_conservative(self):
    return self.conservative_type()(**self.params())

  def _from_flux(self, flux):
    return -staggered_first_derivative(flux, self.grid.solution_dx)


# ---------------------------------------------------------------------------------
# Burgers: u_t + (u^2/2)_x = eta u_xx + forcing          (equations.py:230-370)
# ---------------------------------------------------------------------------------
class BurgersEquation(Equation):
  CONSERVATIVE = False
  GRID_OFFSET = polynomials.GridOffset.CENTERED
  EXACT_METHOD = ExactMethod.WENO
  DERIVATIVE_NAMES = ('u_x', 'u_xx')
  DERIVATIVE_ORDERS = (1, 2)
  KIND = 'burgers'
  FORCED = True
  NUM_FORCING_TERMS = 20

  def __init__(self, num_points, resample_factor=1, period=2 * np.pi, random_seed=0, eta=0.04,
               k_min=1, k_max=3):
    super(BurgersEquation, self).__init__(num_points, resample_factor, period, random_seed,
                                          k_min, k_max)
    self.eta = eta

  def initial_value(self):
    return np.zeros_like(self.grid.solution_x)

  time_step = 1e-3
  standard_deviation = 0.7917

  def equation_of_motion(self, y, spatial_derivatives):
    d = spatial_derivatives
    return self.eta * d['u_xx'] - y * d['u_x']

  def params(self):
    out = super(BurgersEquation, self).params()
    out['eta'] = self.eta
    return out

  @classmethod
  def base_type(cls):
    return BurgersEquation

  @classmethod
  def exact_type(cls):
    return GodunovBurgersEquation


class ConservativeBurgersEquation(BurgersEquation):
  CONSERVATIVE = True
  GRID_OFFSET = polynomials.GridOffset.STAGGERED
  DERIVATIVE_NAMES = ('u', 'u_x')
  DERIVATIVE_ORDERS = (0, 1)
  VARIANT = 'conservative'

  def equation_of_motion(self, y, spatial_derivatives):
    d = spatial_derivatives
    return self._from_flux(0.5 * d['u'] ** 2 - self.eta * d['u_x'])


class GodunovBurgersEquation(BurgersEquation):
  CONSERVATIVE = True
  GRID_OFFSET = polynomials.GridOffset.STAGGERED
  DERIVATIVE_NAMES = ('u_minus', 'u_plus', 'u_x')
  DERIVATIVE_ORDERS = (0, 0, 1)
  VARIANT = 'godunov'

  def equation_of_motion(self, y, spatial_derivatives):
    d = spatial_derivatives
    return self._from_flux(godunov_convective_flux(d['u_minus'], d['u_plus']) - self.eta * d['u_x'])


# ---------------------------------------------------------------------------------
# Korteweg-de Vries: u_t + 6 u u_x + u_xxx = 0           (equations.py:373-478)
# ---------------------------------------------------------------------------------
class KdVEquation(Equation):
  CONSERVATIVE = False
  GRID_OFFSET = polynomials.GridOffset.CENTERED
  EXACT_METHOD = ExactMethod.SPECTRAL
  DERIVATIVE_NAMES = ('u_x', 'u_xxx')
  DERIVATIVE_ORDERS = (1, 3)
  KIND = 'kdv'
  NUM_FORCING_TERMS = 10    # only seeds the initial condition

  def __init__(self, num_points, resample_factor=1, period=32, random_seed=0, k_min=1, k_max=3):
    super(KdVEquation, self).__init__(num_points, resample_factor, period, random_seed, k_min, k_max)

  def initial_value(self):
    return self.forcing(0)

  time_step = 2.5e-5
  standard_deviation = 0.594

  def equation_of_motion(self, y, spatial_derivatives):
    d = spatial_derivatives
    return -6 * y * d['u_x'] - d['u_xxx']

  @classmethod
  def base_type(cls):
    return KdVEquation


class ConservativeKdVEquation(KdVEquation):
  CONSERVATIVE = True
  GRID_OFFSET = polynomials.GridOffset.STAGGERED
  DERIVATIVE_NAMES = ('u', 'u_xx')
  DERIVATIVE_ORDERS = (0, 2)
  VARIANT = 'conservative'

  def equation_of_motion(self, y, spatial_derivatives):
    d = spatial_derivatives
    return self._from_flux(3 * d['u'] ** 2 + d['u_xx'])


class GodunovKdVEquation(KdVEquation):
  CONSERVATIVE = True
  GRID_OFFSET = polynomials.GridOffset.STAGGERED
  DERIVATIVE_NAMES = ('u_minus', 'u_plus', 'u_xx')
  DERIVATIVE_ORDERS = (0, 0, 2)
  VARIANT = 'godunov'

  def equation_of_motion(self, y, spatial_derivatives):
    d = spatial_derivatives
    return self._from_flux(6 * godunov_convective_flux(d['u_minus'], d['u_plus']) + d['u_xx'])


# ---------------------------------------------------------------------------------
# Kuramoto-Sivashinsky: u_t + u u_x + u_xx + u_xxxx = 0  (equations.py:481-587)
# ---------------------------------------------------------------------------------
class KSEquation(Equation):
  CONSERVATIVE = False
  GRID_OFFSET = polynomials.GridOffset.CENTERED
  EXACT_METHOD = ExactMethod.SPECTRAL
  DERIVATIVE_NAMES = ('u_x', 'u_xx', 'u_xxxx')
  DERIVATIVE_ORDERS = (1, 2, 4)
  KIND = 'ks'
  NUM_FORCING_TERMS = 10

  def __init__(self, num_points, resample_factor=1, period=64, random_seed=0, k_min=1, k_max=3):
    super(KSEquation, self).__init__(num_points, resample_factor, period, random_seed, k_min, k_max)

  def initial_value(self):
    return self.forcing(0)

  time_step = 2.5e-5
  standard_deviation = 0.299

  def equation_of_motion(self, y, spatial_derivatives):
    d = spatial_derivatives
    return -y * d['u_x'] - d['u_xxxx'] - d['u_xx']

  @classmethod
  def base_type(cls):
    return KSEquation


class ConservativeKSEquation(KSEquation):
  CONSERVATIVE = True
  GRID_OFFSET = polynomials.GridOffset.STAGGERED
  DERIVATIVE_NAMES = ('u', 'u_x', 'u_xxx')
  DERIVATIVE_ORDERS = (0, 1, 3)
  VARIANT = 'conservative'

  def equation_of_motion(self, y, spatial_derivatives):
    d = spatial_derivatives
    return self._from_flux(0.5 * d['u'] ** 2 + d['u_xxx'] + d['u_x'])


class GodunovKSEquation(KSEquation):
  CONSERVATIVE = True
  GRID_OFFSET = polynomials.GridOffset.STAGGERED
  DERIVATIVE_NAMES = ('u_minus', 'u_plus', 'u_x', 'u_xxx')
  DERIVATIVE_ORDERS = (0, 0, 1, 3)
  VARIANT = 'godunov'

  def equation_of_motion(self, y, spatial_derivatives):
    d = spatial_derivatives
    return self._from_flux(d['u_xxx'] + d['u_x'] + godunov_convective_flux(d['u_minus'], d['u_plus']))


# registries (equations.py:590-606)
EQUATION_TYPES = {'burgers': BurgersEquation, 'kdv': KdVEquation, 'ks': KSEquation}
CONSERVATIVE_EQUATION_TYPES = {'burgers': ConservativeBurgersEquation, 'kdv': ConservativeKdVEquation,
                               'ks': ConservativeKSEquation}
FLUX_EQUATION_TYPES = {'burgers': GodunovBurgersEquation, 'kdv': GodunovKdVEquation,
                       'ks': GodunovKSEquation}


def equation_type_from_hparams(hparams):
  """equations.py:609-626."""
  if not hparams.conservative:
    return EQUATION_TYPES[hparams.equation]
  registry = FLUX_EQUATION_TYPES if hparams.numerical_flux else CONSERVATIVE_EQUATION_TYPES
  return registry[hparams.equation]


def from_hparams(hparams, random_seed=0):
  """(fine equation, coarse equation) for a model's hparams (equations.py:629-662)."""
  kwargs = json.loads(hparams.equation_kwargs)
  fine_points = kwargs.pop('num_points')
  coarse_points, rest = divmod(fine_points, hparams.resample_factor)
  if rest:
    raise ValueError('resample_factor={} does not divide exact_num_points={}'
                     .format(hparams.resample_factor, fine_points))
  coarse = equation_type_from_hparams(hparams)(
      coarse_points, resample_factor=hparams.resample_factor, random_seed=random_seed, **kwargs)
  return coarse.to_fine(), coarse
