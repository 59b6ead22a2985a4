"""The reference's known-answer tests for the hot path, re-expressed against the
oracle without TensorFlow.  Each test cites the reference test it restates
(paths relative to /root/reference/pde_superresolution/).  CPU only."""
import numpy as np
import pytest

from oracle import pde_oracle as O

FD, FV = O.FINITE_DIFFERENCES, O.FINITE_VOLUMES


@pytest.mark.parametrize('padding,center,expected', [
    # layers_test.py:49-62
    (0, True, [0, 1, 2]), (1, True, [2, 0, 1, 2]), (2, True, [2, 0, 1, 2, 0]),
    (3, True, [1, 2, 0, 1, 2, 0]), (4, True, [1, 2, 0, 1, 2, 0, 1]),
    (6, True, [0, 1, 2, 0, 1, 2, 0, 1, 2]), (7, True, [2, 0, 1, 2, 0, 1, 2, 0, 1, 2]),
    (0, False, [0, 1, 2]), (1, False, [0, 1, 2, 0]), (2, False, [0, 1, 2, 0, 1]),
    (3, False, [0, 1, 2, 0, 1, 2]), (5, False, [0, 1, 2, 0, 1, 2, 0, 1]),
])
def test_pad_periodic(padding, center, expected):
  x = np.arange(3)[None, :, None]
  np.testing.assert_equal(O.pad_periodic(x, padding, center)[0, :, 0], expected)


def test_nn_conv1d_periodic():
  # layers_test.py:69-86
  x = np.arange(5.0)[None, :, None]
  for filt, expected in (([0., 1., 0.], x[0, :, 0]), ([0., 1.], x[0, :, 0]),
                         ([.5, .5], [2.0, 0.5, 1.5, 2.5, 3.5])):
    f = np.array(filt)[:, None, None]
    np.testing.assert_allclose(O.nn_conv1d_periodic(x, f, center=True)[0, :, 0], expected)


@pytest.mark.parametrize('grid,order,expected', [
    # polynomials_test.py:36-51 (Wikipedia finite difference table)
    ([-1, 0, 1], 1, [-1 / 2, 0, 1 / 2]), ([-1, 0, 1], 2, [1, -2, 1]),
    ([-2, -1, 0, 1, 2], 2, [-1 / 12, 4 / 3, -5 / 2, 4 / 3, -1 / 12]),
    ([0, 1], 1, [-1, 1]), ([0, 2], 1, [-0.5, 0.5]), ([0, 0.5], 1, [-2, 2]),
    ([0, 1, 2, 3, 4], 4, [1, -4, 6, -4, 1]),
])
def test_finite_difference_coefficients(grid, order, expected):
  np.testing.assert_allclose(O.coefficients(np.array(grid), FD, order), expected, atol=1e-12)


@pytest.mark.parametrize('grid,order,expected', [
    # polynomials_test.py:54-76
    ([-0.5, 0.5], 0, [1 / 2, 1 / 2]), ([-1, 1], 0, [1 / 2, 1 / 2]), ([-1.5, -0.5], 0, [-1 / 2, 3 / 2]),
    ([-0.5, 0.5, 1.5], 0, [1 / 3, 5 / 6, -1 / 6]), ([-0.25, 0.25, 0.75], 0, [1 / 3, 5 / 6, -1 / 6]),
    ([2.5, 1.5, 0.5, -0.5, -1.5], 0, [2 / 60, -13 / 60, 47 / 60, 27 / 60, -3 / 60]),
    ([-0.5, 0.5], 1, [-1, 1]), ([-1, 1], 1, [-1 / 2, 1 / 2]), ([0.5, 1.5, 2.5], 1, [-2, 3, -1]),
    ([-1.5, -0.5, 0.5, 1.5], 1, [1 / 12, -5 / 4, 5 / 4, -1 / 12]),
    ([-.75, -0.25, 0.25, 0.75], 1, [1 / 6, -5 / 2, 5 / 2, -1 / 6]),
])
def test_finite_volume_coefficients(grid, order, expected):
  np.testing.assert_allclose(O.coefficients(np.array(grid), FV, order), expected, atol=1e-12)


@pytest.mark.parametrize('grid,method,order', [
    # polynomials_test.py:88-104
    ([-2, -1, 0, 1, 2], FD, 1), ([-2, -1, 0, 1, 2], FD, 2),
    ([-1.5, -0.5, 0.5, 1.5], FD, 1), ([-1.5, -0.5, 0.5, 1.5], FV, 1),
])
def test_polynomial_accuracy_layer_consistency(grid, method, order):
  a, b = O.constraints(np.array(grid), method, order, 2)
  layer = O.PolynomialAccuracyLayer(np.array(grid), method, order, 2)
  z = np.random.RandomState(0).randn(10, layer.input_size)
  out = layer.bias + z @ layer.nullspace
  np.testing.assert_allclose(out @ a.T - b, 0, atol=1e-7)


def test_polynomial_accuracy_layer_bias_zero_padding():
  # polynomials_test.py:106-114
  layer = O.PolynomialAccuracyLayer(np.array([-1.5, -0.5, 0.5, 1.5]), FD, 0, bias_zero_padding=(0, 1))
  expected = np.concatenate([O.coefficients(np.array([-1.5, -0.5, 0.5]), FD, 0), [0.0]])
  np.testing.assert_allclose(layer.bias, expected)


@pytest.mark.parametrize('order,offset,expected,acc', [
    # polynomials_test.py:116-157
    (0, O.CENTERED, [0], 1), (1, O.CENTERED, [-1, 0, 1], 1), (2, O.CENTERED, [-1, 0, 1], 1),
    (3, O.CENTERED, [-2, -1, 0, 1, 2], 1), (4, O.CENTERED, [-2, -1, 0, 1, 2], 1),
    (0, O.STAGGERED, [-0.5, 0.5], 1), (1, O.STAGGERED, [-0.5, 0.5], 1),
    (2, O.STAGGERED, [-1.5, -0.5, 0.5, 1.5], 1), (3, O.STAGGERED, [-1.5, -0.5, 0.5, 1.5], 1),
    (0, O.CENTERED, [-3, -2, -1, 0, 1, 2, 3], 6), (0, O.STAGGERED, [-2.5, -1.5, -0.5, 0.5, 1.5, 2.5], 6),
])
def test_regular_grid(order, offset, expected, acc):
  np.testing.assert_allclose(O.regular_grid(offset, order, acc), expected)


def test_weno_smooth_limits():
  # weno_test.py:30-46
  u = np.zeros(5)
  np.testing.assert_allclose(O.weno_omega(u), np.stack(5 * [[0.1, 0.6, 0.3]], axis=1))
  # smooth-limit coefficients: reconstruct a delta to read the stencil
  e = np.zeros(9)
  e[4] = 1e-9  # tiny bump keeps the nonlinear weights at their linear values
  left = O.weno_reconstruct_left(e) / 1e-9
  np.testing.assert_allclose(left[[2, 3, 4, 5, 6]], [2 / 60, -13 / 60, 47 / 60, 27 / 60, -3 / 60][::-1],
                             atol=1e-6)


def test_weno_discontinuity():
  # weno_test.py:48-58
  u = np.array([0, 1, 2, 3, 4, -4, -3, -2, -1.])
  np.testing.assert_allclose(O.weno_reconstruct_left(u),
                             [0.5, 1.5, 2.5, 3.5, 4.5, -3.5, -2.5, -1.5, -0.5], atol=0.005)
  np.testing.assert_allclose(O.weno_reconstruct_right(u),
                             [0.5, 1.5, 2.5, 3.5, -4.5, -3.5, -2.5, -1.5, -0.5], atol=0.005)


@pytest.mark.parametrize('u', [
    [0, 0, 0, 0, 1, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 0, 0, 0, 0, 0], [1, 2, 3, 4, 5, 0, 0, 0, 0, 0],
    [0, 0, 1, 2, 3, 0, 0, 0, 0, 0], [0, 0, 0, 1, 2, 0, 0, 0, 0, 0],
    list(2 * np.random.RandomState(0).rand(10)),
])
def test_weno_symmetry(u):
  # weno_test.py:60-83
  u = np.array(u, dtype=float)
  flip = lambda x: x[::-1]
  flip_staggered = lambda x: flip(np.roll(x, +1))
  np.testing.assert_allclose(O.weno_reconstruct_left(u),
                             flip_staggered(O.weno_reconstruct_right(flip(u))), atol=1e-6)
  np.testing.assert_allclose(O.weno_reconstruct_right(u),
                             flip_staggered(O.weno_reconstruct_left(flip(u))), atol=1e-6)


def test_weno_batched():
  # weno_test.py:85-97
  ub = np.array([[0, 0, 0, 1, 2, 3, 4], [0, 0, 1, 2, 3, 4, 5.]])
  np.testing.assert_allclose(O.weno_reconstruct_left(ub),
                             np.stack([O.weno_reconstruct_left(ub[0]), O.weno_reconstruct_left(ub[1])]))
  np.testing.assert_allclose(O.weno_reconstruct_right(ub),
                             np.stack([O.weno_reconstruct_right(ub[0]), O.weno_reconstruct_right(ub[1])]))


def test_resample():
  # duckarray_test.py:32-54
  np.testing.assert_allclose(O.resample_mean(np.arange(6.0), 2), [0.5, 2.5, 4.5])
  np.testing.assert_allclose(O.subsample(np.arange(6), 2), [0, 2, 4])


def test_spectral_derivative_matches_fftpack():
  # duckarray_test.py:56-66
  import scipy.fftpack
  for y, period in ((np.sin(2 * np.pi * np.arange(8) / 8), 1), (np.sin(2 * np.pi * np.arange(8) / 8), 8),
                    (np.linspace(-1, 1, num=12) ** 2, 2)):
    for order in range(3):
      np.testing.assert_allclose(scipy.fftpack.diff(y, order=order, period=period),
                                 O.spectral_derivative(y, order, period), atol=1e-12)


def test_conv_stack_against_torch():
  """Independent second opinion on the conv alignment: torch's circular conv1d."""
  import torch
  import torch.nn.functional as F
  rs = np.random.RandomState(3)
  x = rs.randn(2, 20, 3).astype(np.float32)
  for k in (2, 3, 4, 5):
    w = rs.randn(k, 3, 4).astype(np.float32)
    b = rs.randn(4).astype(np.float32)
    ours = O.conv1d_periodic_layer(x, w, b, 'relu', center=True)
    left, right = -(-(k - 1) // 2), (k - 1) // 2
    xt = torch.from_numpy(x).permute(0, 2, 1)            # [b, c, x]
    xt = torch.cat([xt[..., xt.shape[-1] - left:], xt, xt[..., :right]], dim=-1) if k > 1 else xt
    wt = torch.from_numpy(w).permute(2, 1, 0).contiguous()  # [out, in, k]
    ref = F.relu(F.conv1d(xt, wt, torch.from_numpy(b))).permute(0, 2, 1).numpy()
    np.testing.assert_allclose(ours, ref, rtol=1e-5, atol=1e-5)


def test_burgers_baseline_vs_weno_consistency():
  """integrate_test.py:187-198 (Godunov Burgers, accuracy-1 baseline vs WENO within 1e-3),
  shortened to t<=0.3 to keep the CPU suite fast."""
  eq = O.EquationSpec('burgers', 'godunov', num_points=200)
  times = np.linspace(0, 0.3, 4)
  yb, _ = O.odeint(eq.initial_value(), O.PolynomialDifferentiator(eq, 1), times)
  yw, _ = O.odeint(eq.initial_value(), O.WENODifferentiator(eq), times)
  np.testing.assert_allclose(yb, yw, rtol=1e-3, atol=1e-3)
  assert abs(yw.mean(axis=1)).max() < 1e-3      # integrate_test.py:141-144


def test_fixed_step_matches_scipy_when_pinned():
  """With the controller pinned at max_step the reference's RK23 IS fixed-step
  Bogacki-Shampine: the oracle's fixed-step integrator reproduces C1 at t=2
  up to the two start-up steps SciPy takes (h=1e-4, 1e-3)."""
  eq = O.EquationSpec('burgers', num_points=64)
  rhs = O.batched_rhs([eq], mode='fd', accuracy_order=1)
  y = O.fixed_step_integrate(rhs, eq.initial_value()[None], 0.0, 0.01, 200, save_every=200)[0, 0]
  ys, _ = O.odeint(eq.initial_value(), O.PolynomialDifferentiator(eq, 1), np.array([0.0, 2.0]))
  np.testing.assert_allclose(y, ys[-1], rtol=0, atol=2e-4)
