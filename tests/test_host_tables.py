"""Host-side mirror (ddd1d_b200.polynomials / equations / training) against fixtures
minted from the reference's own NumPy code, and the C-ABI surface.  CPU only: no
kernel is launched."""
import ctypes
import os
import re

import numpy as np
import pytest

import ddd1d_b200 as ddd
from ddd1d_b200 import _lib, equations, polynomials, runtime, training
from tests.helpers import KINDS, VARIANTS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REGISTRY = {'plain': equations.EQUATION_TYPES, 'conservative': equations.CONSERVATIVE_EQUATION_TYPES,
            'godunov': equations.FLUX_EQUATION_TYPES}


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('variant', VARIANTS)
def test_tables_match_reference(golden, kind, variant):
  g = golden('tables')
  hp = training.create_hparams(kind)
  for n in (32, 64, 256):
    eq = REGISTRY[variant][kind](n)
    key = '%s/%s/%d' % (kind, variant, n)
    np.testing.assert_allclose(runtime.coefficient_grid(eq, hp), g[key + '/grid'], rtol=0, atol=1e-15)
    for d, (layer, order) in enumerate(zip(runtime.accuracy_layers(eq, hp), eq.DERIVATIVE_ORDERS)):
      np.testing.assert_allclose(layer.bias, g['%s/bias%d' % (key, d)], rtol=1e-12, atol=1e-12)
      # identical constraint matrix => identical LAPACK null-space basis, sign included
      np.testing.assert_allclose(layer.nullspace, g['%s/nullspace%d' % (key, d)], rtol=1e-10, atol=1e-9)
      for acc in (1, 3):
        grid = polynomials.regular_grid(eq.GRID_OFFSET, order, acc, eq.grid.solution_dx)
        np.testing.assert_allclose(grid, g['%s/fdgrid%d_acc%d' % (key, d, acc)], rtol=0, atol=1e-15)
        np.testing.assert_allclose(polynomials.coefficients(grid, runtime.method_for(eq), order),
                                   g['%s/fdcoef%d_acc%d' % (key, d, acc)], rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('variant', VARIANTS)
def test_equations_match_reference(golden, kind, variant):
  g = golden('pointwise')
  eq = REGISTRY[variant][kind](24, random_seed=2)
  key = '%s/%s' % (kind, variant)
  derivs = {name: g['%s/deriv/%s' % (key, name)] for name in eq.DERIVATIVE_NAMES}
  np.testing.assert_allclose(eq.equation_of_motion(g['y'], derivs), g[key + '/equation_of_motion'],
                             rtol=1e-13, atol=1e-13)
  np.testing.assert_allclose(eq.initial_value(), g[key + '/initial_value'], rtol=1e-13, atol=1e-14)
  assert eq.time_step == float(g[key + '/time_step'])
  assert eq.standard_deviation == float(g[key + '/standard_deviation'])
  # class plumbing (equations.py:160-193)
  assert type(eq.to_fine()) is type(eq)
  assert eq.to_conservative().CONSERVATIVE
  assert eq.params()['num_points'] == 24
  exact = eq.to_exact()
  assert type(exact) is (equations.GodunovBurgersEquation if kind == 'burgers'
                         else equations.EQUATION_TYPES[kind])


def test_forcing_matches_reference(golden):
  g = golden('pointwise')
  for seed in (0, 1, 17):
    for factor, cls in ((1, equations.BurgersEquation), (4, equations.BurgersEquation),
                        (4, equations.ConservativeBurgersEquation)):
      eq = cls(16, resample_factor=factor, random_seed=seed)
      key = 'forcing/%d/%d/%d' % (seed, factor, int(eq.CONSERVATIVE))
      for name in ('a', 'omega', 'k', 'phi'):
        np.testing.assert_array_equal(getattr(eq.forcing, name), g['%s/%s' % (key, name)])
      for t in (0.0, 0.731, 12.5):
        np.testing.assert_allclose(eq.forcing(t), g['%s/t%g/f64' % (key, t)], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(eq.finalize_time_derivative(t, np.zeros(16)),
                                   g['%s/t%g/f64' % (key, t)], rtol=1e-13, atol=1e-14)


def test_from_hparams_and_defaults():
  hp = training.create_hparams('kdv', resample_factor=4, equation_kwargs='{"num_points": 64}')
  assert (hp.num_layers, hp.filter_size, hp.kernel_size, hp.nonlinearity) == (3, 32, 5, 'relu')
  assert hp.conservative and not hp.numerical_flux and hp.coefficient_grid_min_size == 6
  fine, coarse = equations.from_hparams(hp, random_seed=3)
  assert type(coarse) is equations.ConservativeKdVEquation
  assert coarse.grid.solution_num_points == 16 and fine.grid.solution_num_points == 64
  with pytest.raises(ValueError):
    equations.from_hparams(training.create_hparams('kdv', resample_factor=5,
                                                   equation_kwargs='{"num_points": 64}'))
  with pytest.raises(ValueError):
    training.create_hparams('kdv', not_a_hparam=1)


def test_window_alignment():
  # s points sit ceil((s-1)/2) to the left of the output (layers.py:76-79)
  centre = lambda w: list(w[2:9])          # the 7 slots -3..+3 of the 11-slot window
  assert _lib.WINDOW == 11
  for stencil, want in (([1., 2., 3.], [0, 0, 1, 2, 3, 0, 0]), ([1., 2.], [0, 0, 1, 2, 0, 0, 0]),
                        ([1., 2., 3., 4.], [0, 1, 2, 3, 4, 0, 0]), (np.arange(6.) + 1, [1, 2, 3, 4, 5, 6, 0]),
                        (np.arange(7.) + 1, np.arange(7.) + 1)):
    w = _lib.to_window(stencil)
    np.testing.assert_array_equal(centre(w), want)
    assert not w[:2].any() and not w[9:].any()
  # hparams.coefficient_grid_min_size = 9: 9 centred points (-4..+4), 10 staggered ones (-5..+4)
  np.testing.assert_array_equal(_lib.to_window(np.arange(9.) + 1), [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0])
  np.testing.assert_array_equal(_lib.to_window(np.arange(10.) + 1), [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 0])
  with pytest.raises(NotImplementedError):
    _lib.to_window(np.arange(12.))


def test_abi_exports_every_declared_symbol():
  header = open(os.path.join(ROOT, 'include', 'ddd1d.h')).read()
  declared = sorted(set(re.findall(r'\b(ddd1d_[a-z0-9_]+)\s*\(', header)))
  assert declared == _lib.exported_symbols()
  lib = _lib.load()   # raises loudly if the library is missing
  for name in declared:
    assert getattr(lib, name) is not None
  assert lib.ddd1d_version() == 2
  assert ctypes.sizeof(_lib.Config) == 88


def test_debug_library_exports_every_declared_symbol():
  """The test-only library (include/ddd1d_debug.h): laboratory kernels, never loaded by the product path; no symbol
  of it lives in the product library."""
  header = open(os.path.join(ROOT, 'include', 'ddd1d_debug.h')).read()
  declared = sorted(set(re.findall(r'\b(ddd1d_debug_[a-z0-9_]+)\s*\(', header)))
  assert declared == sorted(_lib._DEBUG_SIGNATURES)
  debug = _lib.load_debug()
  product = _lib.load()
  for name in declared:
    assert getattr(debug, name) is not None
    assert not hasattr(product, name)


def test_create_rejects_bad_configs_without_gpu():
  """Argument validation happens before any CUDA call (the reference's ValueErrors)."""
  lib = _lib.load()
  handle = ctypes.c_void_p()
  cfg = _lib.Config()
  cfg.struct_bytes = ctypes.sizeof(_lib.Config)
  cfg.equation, cfg.variant, cfg.mode = _lib.BURGERS, _lib.PLAIN, _lib.MODE_STENCIL
  cfg.num_points, cfg.num_derivatives, cfg.dx = 64, 3, 0.1      # Burgers has 2 channels
  assert lib.ddd1d_create(ctypes.byref(cfg), ctypes.byref(handle)) == _lib.EINVAL
  assert b'derivative channels' in lib.ddd1d_last_error(None)
  cfg.num_derivatives, cfg.mode = 2, _lib.MODE_WENO             # WENO needs a Godunov equation
  assert lib.ddd1d_create(ctypes.byref(cfg), ctypes.byref(handle)) == _lib.EINVAL
  cfg.struct_bytes = 4
  assert lib.ddd1d_create(ctypes.byref(cfg), ctypes.byref(handle)) == _lib.EINVAL


def test_no_oracle_import_in_product():
  """The product package must never import the oracle (or any CPU fallback)."""
  pkg = os.path.join(ROOT, 'data-driven-discretization-1d_b200')
  for name in os.listdir(pkg):
    if name.endswith('.py'):
      text = open(os.path.join(pkg, name)).read()
      assert 'oracle' not in text, name


def test_hparams_set_and_parse():
  """tf.contrib.training.HParams surface used by the scripts: set_hparam / parse with type casting."""
  hp = training.create_hparams('burgers')
  hp.set_hparam('equation_kwargs', '{"num_points": 128}')
  hp.set_hparam('num_layers', 4.0)
  assert hp.equation_kwargs == '{"num_points": 128}' and hp.num_layers == 4 and isinstance(hp.num_layers, int)
  hp.parse('conservative=false, filter_size=16,nonlinearity=tanh,learning_rates=[0.01,0.001],resample_factor=8')
  assert hp.conservative is False and hp.filter_size == 16 and hp.nonlinearity == 'tanh'
  assert hp.learning_rates == [0.01, 0.001] and hp.resample_factor == 8
  with pytest.raises(KeyError):
    hp.set_hparam('nope', 1)
  with pytest.raises(ValueError):
    hp.set_hparam('num_layers', 2.5)
  with pytest.raises(ValueError):
    hp.set_hparam('learning_rates', 0.1)
  with pytest.raises(ValueError):
    hp.parse('filter_size')
