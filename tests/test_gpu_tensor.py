"""The tcgen05 / TMEM engine of the learned-coefficient kernel (csrc/ddd1d_tc.cuh):
descriptor probe, parity against the oracle and against the FFMA engine.  Same
float32 tolerances as the FFMA path (the tensor path is 3xTF32).  Runs after
test_gpu_parity.py (alphabetical) so that a fault here cannot poison those tests."""
import ctypes

import numpy as np
import pytest

from oracle import pde_oracle as O
from tests.helpers import KINDS, VARIANTS, assert_f32_faithful, rel_err
from tests import gpu_helpers as G

pytestmark = pytest.mark.gpu

RHS_TOL = 1e-5
FLUX_TOL = 3e-5
TRAJ_TOL = 1e-4


def cpu(t):
  return t.detach().cpu().numpy()


def _round_tf32(a):
  return ((a.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)


def pack_b_planes(w):
  """[5][32][nout] filters -> B planes [5*8][2*nout][4]: rows 0..nout-1 = TF32(w), the rest = TF32(w - TF32(w))."""
  nout = w.shape[2]
  packed = np.zeros((5 * 8, 2 * nout, 4), np.float32)
  hi = _round_tf32(np.ascontiguousarray(w))
  lo = _round_tf32((w - hi).astype(np.float32))
  for k in range(5):
    for ci in range(32):
      packed[k * 8 + ci // 4, :nout, ci % 4] = hi[k, ci, :]
      packed[k * 8 + ci // 4, nout:, ci % 4] = lo[k, ci, :]
  return packed


@pytest.mark.parametrize('nout', (32, 16))
def test_descriptor_probe(nout):
  """One 128-position tile, 32 -> nout channels, 5 taps: validates the shared-memory
  descriptors (tap shift by start address), the [Whi|Wlo] B packing, the 3xTF32 split, the
  main/cross x even/odd accumulator layout and the TMEM read-back in isolation."""
  import torch
  from ddd1d_b200 import _lib
  lib = _lib.load_debug()
  rs = np.random.RandomState(nout)
  x = rs.randn(132, 32).astype(np.float32)
  w = (rs.randn(5, 32, nout) / 8).astype(np.float32)
  dx, dw = (torch.as_tensor(a).cuda().contiguous() for a in (x, pack_b_planes(w)))
  out = torch.zeros((128, nout), dtype=torch.float32, device='cuda')
  _lib.check(lib.ddd1d_debug_tc_probe(0, dx.data_ptr(), dw.data_ptr(), out.data_ptr(), nout,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
  torch.cuda.synchronize()
  want = np.zeros((128, nout))
  for k in range(5):
    want += x[k:k + 128].astype(np.float64) @ w[k].astype(np.float64)
  err = rel_err(cpu(out), want)
  assert err < 1e-6, err


def _solvers(kind, variant, n, seed=0):
  from ddd1d_b200 import runtime
  eq = G.product_equation(kind, variant, n, seed=3)
  hp = G.product_hparams(kind, variant, n)
  oeq = G.oracle_equation(kind, variant, n, seed=3)
  w = O.glorot_weights(oeq, O.NetSpec(), seed=seed, last_layer_scale=0.1, bias_scale=0.1)
  tensor = runtime.learned_solver(eq, hp, w, engine='tensor')
  ffma = runtime.learned_solver(eq, hp, w, engine='ffma')
  assert tensor.engine() == 'tensor' and ffma.engine() == 'ffma'
  return tensor, ffma, oeq, w


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('variant', VARIANTS)
def test_tensor_engine_per_call_parity(kind, variant):
  """Coefficients within RHS_TOL of the float32 oracle; derivatives and dy/dt (where 1/dx^n
  stencils and flux differences amplify rounding) as accurate as the float32 reference graph
  itself, measured against the float64 oracle; and the same bound for the FFMA engine."""
  n = 128
  tensor, ffma, oeq, w = _solvers(kind, variant, n)
  u = G.smooth_rows(5, n, seed=7)
  net = O.NetSpec()
  c32 = O.predict_coefficients(u, oeq, net, w)
  c64 = O.predict_coefficients(u, oeq, net, w, dtype=np.float64)
  assert rel_err(cpu(tensor.coefficients(u)), c32) < RHS_TOL
  assert_f32_faithful(cpu(tensor.coefficients(u)), c32, c64, what='coefficients')
  d32, d64 = O.apply_coefficients(c32, u), O.apply_coefficients(c64, u.astype(np.float64))
  assert_f32_faithful(cpu(tensor.space_derivatives(u)), d32, d64, what='space derivatives (tensor)')
  assert_f32_faithful(cpu(ffma.space_derivatives(u)), d32, d64, what='space derivatives (ffma)')
  t32 = oeq.finalize_time_derivative(np.float32(0.4), O.predict_time_derivative(u[:1], oeq, net, w),
                                     dtype=np.float32)
  t64 = oeq.finalize_time_derivative(0.4, O.predict_time_derivative(u[:1], oeq, net, w, dtype=np.float64))
  assert_f32_faithful(cpu(tensor.rhs(0.4, u[:1])), t32, t64, what='dy/dt (tensor)')
  assert_f32_faithful(cpu(ffma.rhs(0.4, u[:1])), t32, t64, what='dy/dt (ffma)')
  assert rel_err(cpu(tensor.rhs(0.4, u[:1])), cpu(ffma.rhs(0.4, u[:1]))) < 3e-4      # sanity: same quantity
  # float64 rows (what SciPy hands over)
  assert_f32_faithful(cpu(tensor.rhs(0.4, u[:1].astype(np.float64))), t32, t64, what='dy/dt from float64 rows')


@pytest.mark.parametrize('n,kind,dt', ((256, 'burgers', 1e-3), (128, 'kdv', 2.5e-5), (512, 'ks', 1e-5)))
def test_tensor_engine_trajectories(n, kind, dt):
  """BASELINE shapes C2/C3/C4 on the tensor engine: odd batch sizes exercise teams that
  own different numbers of rows (and none at all)."""
  import torch
  from ddd1d_b200 import runtime
  for batch in (1, 3, 301):
    eqs = [G.product_equation(kind, 'plain', n, seed=s) for s in range(batch)]
    oeq0 = G.oracle_equation(kind, 'plain', n)
    w = O.glorot_weights(oeq0, O.NetSpec(), seed=1, last_layer_scale=0.01)
    solver = runtime.learned_solver(eqs, G.product_hparams(kind, 'plain', n), w, engine='tensor')
    u0 = G.smooth_rows(batch, n, seed=batch)
    steps = 6
    got, bad = solver.integrate(u0, 0.05, dt, steps, 3, return_first_bad=True)
    assert tuple(got.shape) == (2, batch, n)
    assert (cpu(bad) == -1).all()
    pick = sorted({0, batch // 2, batch - 1})
    oeqs = [G.oracle_equation(kind, 'plain', n, seed=i) for i in pick]
    rhs = O.batched_rhs(oeqs, O.NetSpec(), w, mode='learned')
    want = O.fixed_step_integrate(rhs, u0[pick], 0.05, dt, steps, 3)
    assert rel_err(cpu(got[:, pick]), want) < TRAJ_TOL
    again = solver.integrate(u0, 0.05, dt, steps, 3)
    assert torch.equal(again, got)


def test_tensor_engine_other_nets_and_schemes():
  """2-layer net, raw (accuracy-order 0) projections; midpoint scheme."""
  from ddd1d_b200 import runtime
  n = 128
  for overrides in (dict(num_layers=2), dict(polynomial_accuracy_order=0),
                    dict(polynomial_accuracy_order=0, ensure_unbiased_coefficients=True)):
    eq = G.product_equation('burgers', 'plain', n)
    hp = G.product_hparams('burgers', 'plain', n, **overrides)
    net = O.NetSpec(**overrides)
    oeq = G.oracle_equation('burgers', 'plain', n)
    w = O.glorot_weights(oeq, net, seed=4, last_layer_scale=0.1, bias_scale=0.1)
    solver = runtime.learned_solver(eq, hp, w, engine='tensor', forcing=False)
    u = G.smooth_rows(3, n, seed=1)
    assert rel_err(cpu(solver.coefficients(u)), O.predict_coefficients(u, oeq, net, w)) < RHS_TOL, overrides
    assert_f32_faithful(cpu(solver.rhs(0.0, u)), O.predict_time_derivative(u, oeq, net, w),
                        O.predict_time_derivative(u, oeq, net, w, dtype=np.float64), what=str(overrides))
  snaps = solver.integrate(u, 0.0, 1e-3, 4, 2, 'midpoint')
  rhs = O.batched_rhs([oeq], net, w, mode='learned')
  # forcing disabled on the GPU side: compare with an unforced oracle RHS
  unforced = lambda t, y: O.predict_time_derivative(np.asarray(y, np.float32), oeq, net, w)
  want = O.fixed_step_integrate(unforced, u, 0.0, 1e-3, 4, 2, scheme='midpoint')
  assert rel_err(cpu(snaps), want) < TRAJ_TOL


def test_tensor_engine_rejects_unsupported_shapes():
  from ddd1d_b200 import runtime
  eq = G.product_equation('burgers', 'plain', 200)       # the reference's own test size: not a tile multiple
  hp = G.product_hparams('burgers', 'plain', 200)
  w = O.glorot_weights(G.oracle_equation('burgers', 'plain', 200), O.NetSpec(), seed=0)
  with pytest.raises(NotImplementedError):
    runtime.learned_solver(eq, hp, w, engine='tensor').engine()
  assert runtime.learned_solver(eq, hp, w).engine() == 'ffma'
  # a tanh net at a supported size also takes the FFMA engine
  eq = G.product_equation('burgers', 'plain', 128)
  hp = G.product_hparams('burgers', 'plain', 128, nonlinearity='tanh')
  w = O.glorot_weights(G.oracle_equation('burgers', 'plain', 128), O.NetSpec(nonlinearity='tanh'), seed=0)
  assert runtime.learned_solver(eq, hp, w).engine() == 'ffma'


@pytest.mark.parametrize('n', (32, 64))
@pytest.mark.parametrize('kind', KINDS)
def test_packed_rows_trajectories(n, kind):
  """The reference's own grids (notebooks/time-integration.ipynb: N = 32 / 64): 4 or 2 rows share one
  128-position MMA tile.  Batches that do not fill a tile, a slot or a wave; every row against the oracle;
  conservative form too (its flux difference wraps inside the tile)."""
  import torch
  from ddd1d_b200 import runtime
  dt = {'burgers': 1e-3, 'kdv': 2.5e-5, 'ks': 1e-5}[kind]
  for variant, batch in (('plain', 1), ('plain', 7), ('conservative', 5), ('plain', 2500)):
    eqs = [G.product_equation(kind, variant, n, seed=s) for s in range(batch)]
    oeq0 = G.oracle_equation(kind, variant, n)
    w = O.glorot_weights(oeq0, O.NetSpec(), seed=2, last_layer_scale=0.05, bias_scale=0.1)
    solver = runtime.learned_solver(eqs, G.product_hparams(kind, variant, n), w, engine='tensor')
    assert solver.engine() == 'tensor'
    u0 = G.smooth_rows(batch, n, seed=batch + n)
    steps = 6
    got, bad = solver.integrate(u0, 0.05, dt, steps, 3, return_first_bad=True)
    assert tuple(got.shape) == (2, batch, n)
    assert (cpu(bad) == -1).all()
    pick = sorted({0, 1, batch // 2, batch - 2, batch - 1} & set(range(batch)))
    oeqs = [G.oracle_equation(kind, variant, n, seed=i) for i in pick]
    rhs = O.batched_rhs(oeqs, O.NetSpec(), w, mode='learned')
    want = O.fixed_step_integrate(rhs, u0[pick], 0.05, dt, steps, 3)
    assert rel_err(cpu(got[:, pick]), want) < TRAJ_TOL, (variant, batch)
    assert torch.equal(solver.integrate(u0, 0.05, dt, steps, 3), got)
    # per-call hooks on packed rows
    net = O.NetSpec()
    c32 = O.predict_coefficients(u0[pick], oeq0, net, w)
    assert rel_err(cpu(solver.coefficients(u0))[pick], c32) < RHS_TOL
    solver.close()


@pytest.mark.parametrize('engine,tol', (('tensor_f16x2', 2e-3), ('tensor_f16', 4e-3)))
def test_reduced_precision_engines(engine, tol):
  """The opt-in cheaper operand formats: same kernel, fewer products.  Their per-call accuracy is that of
  fp16 activations (2^-12 relative per element), stated here; the trajectory-level numbers are in
  profiles/r02/tc_trajectory_error.json."""
  from ddd1d_b200 import runtime
  for n in (64, 128, 256):
    eq = G.product_equation('burgers', 'plain', n, seed=3)
    oeq = G.oracle_equation('burgers', 'plain', n, seed=3)
    w = O.glorot_weights(oeq, O.NetSpec(), seed=0, last_layer_scale=0.1, bias_scale=0.1)
    solver = runtime.learned_solver(eq, G.product_hparams('burgers', 'plain', n), w, engine=engine)
    assert solver.engine() == engine
    u = G.smooth_rows(5, n, seed=7)
    c64 = O.predict_coefficients(u, oeq, O.NetSpec(), w, dtype=np.float64)
    err = rel_err(cpu(solver.coefficients(u)), c64)
    assert err < tol, (n, err)
    faithful = runtime.learned_solver(eq, G.product_hparams('burgers', 'plain', n), w, engine='tensor')
    assert rel_err(cpu(faithful.coefficients(u)), c64) < err      # and the faithful form is closer
    solver.close(), faithful.close()


def test_block_pool_three_slots_is_bit_equal(monkeypatch):
  """DDD1D_TC_SLOTS=3 (opt-in): three rows per team on the TMEM accumulator block pool.  Same arithmetic, another
  schedule: results must be bit-identical to the two-slot kernel, for batch sizes that leave teams with three, two,
  one and no rows in their last round, per call and over a trajectory with snapshots."""
  import torch
  from ddd1d_b200 import runtime
  n, kind, dt = 256, 'burgers', 1e-3
  oeq0 = G.oracle_equation(kind, 'plain', n)
  w = O.glorot_weights(oeq0, O.NetSpec(), seed=1, last_layer_scale=0.01)
  sms = torch.cuda.get_device_properties(0).multi_processor_count
  for batch in (2, 2 * sms * 3 + 2 * sms + 5, 2 * sms * 2 - 3):
    eqs = [G.product_equation(kind, 'plain', n, seed=s) for s in range(batch)]
    u0 = G.smooth_rows(batch, n, seed=batch)
    out = {}
    for slots in (2, 3):
      if slots == 3:
        monkeypatch.setenv('DDD1D_TC_SLOTS', '3')
      else:
        monkeypatch.delenv('DDD1D_TC_SLOTS', raising=False)
      solver = runtime.learned_solver(eqs, G.product_hparams(kind, 'plain', n), w, engine='tensor')
      shape = solver.launch_shape(batch)
      assert shape['shared_bytes'] == (230656 if slots == 3 else 164096), shape
      snaps, bad = solver.integrate(u0, 0.05, dt, 7, 3, return_first_bad=True)
      out[slots] = (snaps.clone(), bad.clone(), solver.rhs(0.3, u0).clone())
      solver.close()
    monkeypatch.delenv('DDD1D_TC_SLOTS', raising=False)
    for a, b in zip(out[2], out[3]):
      assert torch.equal(a, b)
