"""The TMEM-resident A operand of tcgen05 (test-only library, include/ddd1d_debug.h): what tcgen05.cp,
tcgen05.shift.down, tcgen05.cp.4x256b and tcgen05.mma with A in TMEM (+ .ashift) do on this hardware.  These are the
measured facts behind DESIGN 4.1's "A operand resident in TMEM" entry (an im2col-free convolution that was probed,
timed and not built); the tests pin them so that the entry stays checkable."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dump():
  from ddd1d_b200 import _lib
  lib = _lib.load_debug()
  out = np.zeros((10, 128, 16), np.uint32)
  _lib.check(lib.ddd1d_debug_tc_shift_probe(0, _lib.host_ptr(out)))
  return out


def _positions(dump, step, col=0):
  """Position index encoded in the low half of a column (the probe's planes hold p | k << 8)."""
  return (dump[step, :, col] & 0xff).astype(int)


def test_cp_128x256b_lays_rows_on_lanes(dump):
  # lane = row of the K-major tile, 8 columns = 16 halfs of K; the copy started at position 2
  np.testing.assert_array_equal(_positions(dump, 0), np.arange(2, 130))
  k_of_col = (dump[0, 5, :8] & 0xffff) >> 8
  np.testing.assert_array_equal(k_of_col, np.arange(0, 16, 2))          # two halfs per 32-bit column, K ascending


def test_shift_down_moves_rows_to_the_next_lower_lane_inside_a_quadrant(dump):
  want = np.arange(3, 131)
  want[[31, 63, 95, 127]] = [33, 65, 97, 129]       # the last lane of every 32-lane quadrant keeps its row
  np.testing.assert_array_equal(_positions(dump, 1), want)


def test_cp_4x256b_writes_one_lane_per_quadrant(dump):
  after_shift, patched = _positions(dump, 1), _positions(dump, 2)
  np.testing.assert_array_equal(patched[[0, 32, 64, 96]], [100, 101, 102, 103])      # four consecutive rows
  rest = np.setdiff1d(np.arange(128), [0, 32, 64, 96])
  np.testing.assert_array_equal(patched[rest], after_shift[rest])
  # the lane field of the address selects the lane inside the quadrant: 31 = the lane a shift leaves stale
  np.testing.assert_array_equal(_positions(dump, 4)[[31, 63, 95, 127]], [140, 141, 142, 143])
  np.testing.assert_array_equal(_positions(dump, 4, col=8)[[16, 48, 80, 112]], [150, 151, 152, 153])


def test_mma_reads_a_from_tmem_and_ashift_shifts_after_use(dump):
  p, k, n = np.arange(160)[:, None], np.arange(16)[None, :], np.arange(16)[:, None]
  a = (((p * 3 + k) % 7) - 3).astype(np.float64)
  b = (((n + 2 * k) % 5) - 2).astype(np.float64)
  full = a @ b.T                                          # D row of the A row at position p
  lanes = np.arange(128)
  np.testing.assert_array_equal(dump[5].view(np.float32), full[lanes])          # plain MMA, A in TMEM
  np.testing.assert_array_equal(dump[6].view(np.float32), full[lanes])          # .ashift: this MMA still sees A unshifted
  shifted = lanes + 1
  shifted[[31, 63, 95, 127]] -= 1
  np.testing.assert_array_equal(dump[7].view(np.float32), full[shifted])        # ... the next one sees it shifted


def test_cp_after_mma_respects_write_after_read():
  from ddd1d_b200 import _lib
  lib = _lib.load_debug()
  for chain in (1, 16, 64):
    out = np.zeros(2 * 128 * 16 + 1, np.uint32)
    _lib.check(lib.ddd1d_debug_tc_war_probe(0, chain, 50, _lib.host_ptr(out)))
    assert int(out[-1]) == 0, (chain, int(out[-1]))
