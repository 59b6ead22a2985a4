"""Stand-alone helpers mirrored from layers.py / weno.py / duckarray.py, on CPU tensors, against
fixtures minted from the reference's own code (tests/golden/make_golden.py: golden_weno_parts) and the
reference's KATs (layers_test.py:49-67, weno_test.py:30-46)."""
import numpy as np
import pytest

from ddd1d_b200 import duckarray, layers, weno


@pytest.mark.parametrize('padding,center,expected', [
    (0, True, [0, 1, 2]), (1, True, [2, 0, 1, 2]), (2, True, [2, 0, 1, 2, 0]), (3, True, [1, 2, 0, 1, 2, 0]),
    (4, True, [1, 2, 0, 1, 2, 0, 1]), (6, True, [0, 1, 2, 0, 1, 2, 0, 1, 2]),
    (7, True, [2, 0, 1, 2, 0, 1, 2, 0, 1, 2]), (0, False, [0, 1, 2]), (1, False, [0, 1, 2, 0]),
    (2, False, [0, 1, 2, 0, 1]), (3, False, [0, 1, 2, 0, 1, 2]), (5, False, [0, 1, 2, 0, 1, 2, 0, 1]),
])
def test_pad_periodic_reference_kats(padding, center, expected):
  out = layers.pad_periodic(np.arange(3)[None, :, None], padding, center=center)
  np.testing.assert_array_equal(out[0, :, 0].numpy(), expected)


def test_pad_periodic_against_reference_fixture(golden):
  g = golden('weno_parts')
  for padding in (0, 1, 2, 3, 4, 6, 11, 13):
    for center in (False, True):
      got = layers.pad_periodic(g['pad/x'], padding, center=center).numpy()
      np.testing.assert_array_equal(got, g['pad/%d/%d' % (padding, int(center))])
  with pytest.raises(ValueError):
    layers.pad_periodic(np.zeros((3, 4)), 2)


def test_weno_parts_against_reference_fixture(golden):
  g = golden('weno_parts')
  u = g['u']
  np.testing.assert_allclose(weno.calculate_smoothness_indicators(u), g['indicators'], rtol=1e-13, atol=1e-13)
  np.testing.assert_allclose(weno.calculate_omega(u), g['omega'], rtol=1e-12, atol=1e-13)
  np.testing.assert_allclose(weno.calculate_omega(u, weno.OPTIMAL_SMOOTH_WEIGHTS[::-1]), g['omega_reversed'],
                             rtol=1e-12, atol=1e-13)
  np.testing.assert_allclose(weno.left_coefficients(u), g['left_coefficients'], rtol=1e-12, atol=1e-13)
  np.testing.assert_allclose(weno.right_coefficients(u), g['right_coefficients'], rtol=1e-12, atol=1e-13)
  # coefficients x shifted rows reproduce the reference's reconstructions (weno.py:92-97,118-123)
  left = sum(weno.left_coefficients(u)[..., i] * np.roll(u, s, axis=-1) for i, s in enumerate([2, 1, 0, -1, -2]))
  right = sum(weno.right_coefficients(u)[..., i] * np.roll(u, s, axis=-1) for i, s in enumerate([1, 0, -1, -2, -3]))
  np.testing.assert_allclose(left, g['left'], rtol=1e-12, atol=1e-12)
  np.testing.assert_allclose(right, g['right'], rtol=1e-12, atol=1e-12)


def test_weno_smooth_limit_kats():
  # weno_test.py:30-46
  u = np.sin(np.linspace(0, 2 * np.pi, 1000, endpoint=False))
  np.testing.assert_allclose(weno.calculate_omega(u)[:, 500], [0.1, 0.6, 0.3], atol=1e-3)
  np.testing.assert_allclose(weno.left_coefficients(u)[500], [2 / 60, -13 / 60, 47 / 60, 27 / 60, -3 / 60], atol=1e-3)
  np.testing.assert_allclose(weno.right_coefficients(u)[500], [-3 / 60, 27 / 60, 47 / 60, -13 / 60, 2 / 60], atol=1e-3)


def test_duckarray_spectral_helpers():
  import scipy.fftpack
  import torch
  grid = np.linspace(0, 7.0, 32, endpoint=False)
  x = np.stack([np.sin(2 * np.pi * m * grid / 7.0 + m) + 0.3 * np.cos(2 * np.pi * (m + 2) * grid / 7.0) for m in (1, 2, 3)])
  for order in (1, 2, 3):
    want = np.stack([scipy.fftpack.diff(row, order=order, period=7.0) for row in x])   # duckarray_test.py:56-66
    np.testing.assert_allclose(duckarray.spectral_derivative(x, order, 7.0), want, atol=1e-10)
    np.testing.assert_allclose(duckarray.spectral_derivative(torch.as_tensor(x), order, 7.0).numpy(), want, atol=1e-10)
  f = duckarray.smoothing_filter(x)
  np.testing.assert_allclose(duckarray.smoothing_filter(torch.as_tensor(x)).numpy(), f, atol=1e-12)
  assert np.abs(np.fft.rfft(f)[..., -1]).max() < 1e-12 and np.allclose(np.fft.rfft(f)[..., 0], np.fft.rfft(x)[..., 0])
  with pytest.raises(ValueError):
    duckarray.spectral_derivative(np.zeros(5))
  assert duckarray.get_shape(x) == (3, 32) and duckarray.sum(x, axis=-1, keepdims=True).shape == (3, 1)
