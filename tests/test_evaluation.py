"""Evaluation metrics (analysis.py, scripts/run_evaluation.py:189-210) as torch reductions: the
reference's own KATs (analysis_test.py:31-42) plus NumPy restatements of the xarray expressions."""
import numpy as np
import pytest

from ddd1d_b200 import evaluation as E


@pytest.mark.parametrize('data,expected', [
    (np.arange(100) < 65, 6.5),
    (np.ones(100), 9.9),
    (np.zeros(100), 0),
    (np.concatenate([np.ones(10), np.zeros(1), np.ones(9), np.zeros(80)]), 1),
])
def test_calculate_survival_reference_kats(data, expected):
  assert E.calculate_survival(data, np.arange(100) / 10).item() == expected


def test_calculate_survival_batched():
  good = np.stack([np.arange(100) < 65, np.ones(100, bool), np.zeros(100, bool)])
  np.testing.assert_array_equal(E.calculate_survival(good, np.arange(100) / 10), [6.5, 9.9, 0.0])


def test_unify_and_mae():
  rs = np.random.RandomState(0)
  samples, times, n, factor = 3, 11, 16, 4
  exact_high = rs.randn(samples, times, n * factor)
  model = rs.randn(samples, times, n)
  t = np.linspace(0, 1, times)
  exact_low = E.unify_x_coords(exact_high, factor).numpy()
  np.testing.assert_allclose(exact_low, exact_high.reshape(samples, times, n, factor).mean(-1), rtol=1e-12)
  stop = [0.35, 1.0, 5.0]
  mae = E.calculate_mae(model, exact_low, t, stop)
  assert mae.shape == (3, samples)
  for i, tm in enumerate(stop):
    sel = t <= tm                                        # xarray label slice is inclusive
    np.testing.assert_allclose(mae[i], np.abs(model - exact_low)[:, sel].mean(axis=(1, 2)), rtol=1e-12)
  model[1, 3, 2] = np.nan                                # skipna=False: NaN poisons every slice containing it
  mae = E.calculate_mae(model, exact_low, t, stop)
  assert np.isnan(mae[:, 1]).tolist() == [True, True, True] and not np.isnan(mae[:, 0]).any()


def test_mostly_good_survival_matches_numpy():
  rs = np.random.RandomState(1)
  samples, times, n, factor = 4, 21, 8, 2
  t = np.arange(times) * 0.5
  exact_high = rs.randn(samples, times, n * factor)
  exact_low = exact_high.reshape(samples, times, n, factor).mean(-1)
  model = exact_low + 0.02 * t[None, :, None] * rs.randn(samples, times, n)      # error grows in time
  q = 0.8
  got = E.mostly_good_survival(model, exact_high, t, quantile=q)
  max_error = np.quantile(np.abs(exact_high), 1 - q)
  good = (np.abs(model - exact_low) <= max_error).mean(-1) >= q
  want = np.where(good.all(1), t.max(), t[np.argmin(good, axis=1)])
  np.testing.assert_array_equal(got, want)
  assert (got < t.max()).any() and (got > 0).any()       # the case is not degenerate
  # NaNs in the exact solution are skipped by the quantile (DataArray.quantile = nanpercentile)
  holed = exact_high.copy()
  holed[0, -1, :3] = np.nan
  got = E.mostly_good_survival(model, holed, t, quantile=q)
  max_error = np.nanquantile(np.abs(holed), 1 - q)
  low = holed.reshape(samples, times, n, factor).mean(-1)
  good = (np.abs(model - low) <= max_error).mean(-1) >= q
  np.testing.assert_array_equal(got, np.where(good.all(1), t.max(), t[np.argmin(good, axis=1)]))


@pytest.mark.parametrize('name', ('results.nc', 'results.npz'))
def test_results_round_trip(tmp_path, name):
  res = {'y': np.random.RandomState(2).randn(2, 3, 8), 'time': np.array([10.0, 10.5, 11.0]),
         'x': np.arange(8) * 0.1, 'num_evals': np.array([100, 103]), 'sample': np.array([0, 1])}
  res['y'][1, 2:] = np.nan                               # a diverged sample is NaN padded (integrate.py:161-167)
  path = str(tmp_path / name)
  E.write_results(path, res)
  back = E.read_results(path)
  for k in res:
    np.testing.assert_array_equal(back[k], res[k])
  with pytest.raises(ValueError):
    E.write_results(path, dict(res, y=res['y'][0]))


def test_results_nc_is_netcdf3_with_the_reference_schema(tmp_path):
  """scripts/run_evaluation.py:168-174: {y: (sample, time, x)}, coords time, x, sample, num_evals(sample), as the
  NetCDF-3 bytes Dataset.to_netcdf() yields (xarray_beam.py:32-35); read back with SciPy's own reader."""
  from scipy.io import netcdf_file
  res = {'y': np.random.RandomState(3).randn(4, 5, 16), 'time': np.linspace(0, 1, 5), 'x': np.arange(16) * 0.5,
         'num_evals': np.array([31, 32, 33, 34]), 'sample': np.arange(4) + 10}
  path = str(tmp_path / 'results.nc')
  E.write_results(path, res)
  with open(path, 'rb') as f:
    assert f.read(3) == b'CDF'                            # the classic format, not HDF5
  with netcdf_file(path, 'r', mmap=False) as f:
    assert {k: v for k, v in f.dimensions.items()} == {'sample': 4, 'time': 5, 'x': 16}
    assert f.variables['y'].dimensions == ('sample', 'time', 'x')
    assert f.variables['num_evals'].dimensions == ('sample',)
    assert f.variables['y'].coordinates == b'num_evals'
    np.testing.assert_array_equal(f.variables['y'][...], res['y'])
    np.testing.assert_array_equal(f.variables['sample'][...], res['sample'])
  mae = np.abs(np.random.RandomState(4).randn(3, 4))
  E.write_metric(str(tmp_path / 'mae.nc'), 'mae', mae, stop_times=[5, 10, 15], samples=res['sample'])
  E.write_metric(str(tmp_path / 'survival.nc'), 'survival', mae[0], samples=res['sample'])
  with netcdf_file(str(tmp_path / 'mae.nc'), 'r', mmap=False) as f:
    assert f.variables['mae'].dimensions == ('time_max', 'sample')
    np.testing.assert_array_equal(f.variables['mae'][...], mae)
    np.testing.assert_array_equal(f.variables['time_max'][...], [5, 10, 15])
  with netcdf_file(str(tmp_path / 'survival.nc'), 'r', mmap=False) as f:
    np.testing.assert_array_equal(f.variables['survival'][...], mae[0])
