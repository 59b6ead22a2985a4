"""Trajectory-level evidence for the tensor engine's operand precision ("the choice must be
made on measured trajectory error"): relative L-inf error against the float64 oracle over the CONFIGURED
horizon -- BASELINE configs 2-4, 10 000 Bogacki-Shampine RK3 steps, rows 0 / 1 / 2047 / 4095 of the batch, the
trajectories of tests/golden/long_horizon.npz -- for every operand scheme:

  emulated on the CPU (`--emulate`, this container; NumPy oracle with the tensor layers' operands rounded):
      tf32      both operands rounded to TF32 (10-bit mantissa), one product       [no kernel: 2x the operand bytes of f16]
      f16       both operands rounded to 11 bits, one product                      (= DDD1D_ENGINE_TENSOR_F16)
      f16x2     activations 11 bits, filters 22 bits                               (= DDD1D_ENGINE_TENSOR_F16X2)
      f16x3     both operands 22 bits (hi*Wh + hi*Wl + lo*Wh)                       (= DDD1D_ENGINE_TENSOR)
  measured on the GPU (`--gpu`): the four engines ffma / tensor / tensor_f16x2 / tensor_f16 as built.

Test infrastructure (imports oracle/).  Each run merges its rows into the JSON given by --out:
  {config: {scheme: {"vs_float64": worst over snapshots and rows, "vs_float32_oracle": ..., "final_vs_float64": ...}},
   "float32_oracle": {config: its own drift vs float64}}
"""
import argparse
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CONFIGS = ('c2', 'c3', 'c4')
SAVE_EVERY = 1000


def _round_bits(a, bits, truncate=False):
  """Round float32 values to `bits` significant bits (round to nearest, ties away): the relative rounding of a
  value scaled into fp16's normal range (bits = 11) or of a hi + lo pair (22).  truncate=True drops the low bits
  instead, which is what the tensor core does to float32 inputs of a kind::tf32 MMA."""
  a = np.ascontiguousarray(a, dtype=np.float32)
  drop = 24 - bits
  if drop <= 0:
    return a
  u = a.view(np.uint32)
  half = np.uint32(0 if truncate else 1 << (drop - 1))
  mask = np.uint32(0xffffffff ^ ((1 << drop) - 1))
  return ((u + half) & mask).view(np.float32)


SCHEMES = {'tf32': (11, 11, True), 'f16': (11, 11, False), 'f16x2': (11, 22, False),
           'f16x3': (22, 22, False)}      # (activation bits, filter bits, truncating)


def _errors(got, want32, want64):
  worst64 = worst32 = 0.0
  for i in range(got.shape[0]):
    for r in range(got.shape[1]):
      scale = np.abs(want64[i, r]).max()
      worst64 = max(worst64, np.abs(got[i, r] - want64[i, r]).max() / scale)
      worst32 = max(worst32, np.abs(got[i, r] - want32[i, r]).max() / scale)
  final = max(np.abs(got[-1, r] - want64[-1, r]).max() / np.abs(want64[-1, r]).max() for r in range(got.shape[1]))
  return {'vs_float64': float(worst64), 'vs_float32_oracle': float(worst32), 'final_vs_float64': float(final)}


def _emulate(args):
  config, scheme = args
  import ddd1d_b200.workloads as wl
  from oracle import pde_oracle as O
  try:
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=1)
  except ImportError:
    pass
  kind, variant, n, batch, dt, mode = wl.WORKLOADS[config]
  picks = np.asarray(wl.HORIZON_PICKS)
  u0 = wl.horizon_rows(config)[picks]
  eqs = [O.EquationSpec(kind, variant, num_points=n, random_seed=int(s)) for s in picks]
  net, weights = O.NetSpec(), wl.synthetic_weights(kind)
  abits, wbits, trunc = SCHEMES[scheme]
  rounded = [weights[0]] + [(_round_bits(w, wbits, trunc), b) for w, b in weights[1:]]     # the first layer runs in FP32
  plain = O.conv1d_periodic_layer
  calls = {'n': 0}

  def layer(inputs, kernel, bias, activation=None, center=True):
    # the layers after the first take their activations from the tensor pipe's operand format
    i = calls['n'] % len(weights)
    calls['n'] += 1
    if i > 0:
      inputs = _round_bits(inputs, abits, trunc)
      out = plain(inputs.astype(np.float64), kernel.astype(np.float64), bias.astype(np.float64), activation, center)
      return out.astype(np.float32)               # FP32 accumulate ~ exact product sum, one rounding
    return plain(inputs, kernel, bias, activation, center)

  O.conv1d_periodic_layer = layer

  def rhs(t, y):
    y_t = O.predict_time_derivative(np.asarray(y, dtype=np.float32), eqs[0], net, rounded)
    if kind == 'burgers':
      y_t = y_t + np.stack([e.forcing(np.float32(t), dtype=np.float32) for e in eqs])
    return y_t

  out = O.fixed_step_integrate(rhs, u0, 0.0, dt, wl.FULL_STEPS, SAVE_EVERY)
  return config, scheme, out


def run_emulation(out_path):
  fixture = np.load(os.path.join(ROOT, 'tests', 'golden', 'long_horizon.npz'))
  jobs = [(c, s) for c in CONFIGS for s in SCHEMES]
  results = {}
  with mp.get_context('spawn').Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
    for config, scheme, out in pool.imap_unordered(_emulate, jobs):
      e = _errors(out, fixture[config + '/f32'], fixture[config + '/f64'])
      e['how'] = 'NumPy oracle, operands of the tensor layers rounded (emulation)'
      results.setdefault(config, {})['emulated_' + scheme] = e
      print(config, scheme, e, flush=True)
  merge(out_path, results, fixture)


def run_gpu(out_path):
  import ddd1d_b200.workloads as wl
  from ddd1d_b200 import runtime
  from tests import gpu_helpers as G
  fixture = np.load(os.path.join(ROOT, 'tests', 'golden', 'long_horizon.npz'))
  results = {}
  for config in CONFIGS:
    kind, variant, n, batch, dt, mode = wl.WORKLOADS[config]
    picks = fixture[config + '/rows']
    for engine in ('ffma', 'tensor', 'tensor_f16x2', 'tensor_f16'):
      eqs = [G.product_equation(kind, variant, n, seed=int(s)) for s in picks]
      solver = runtime.learned_solver(eqs, G.product_hparams(kind, variant, n), wl.synthetic_weights(kind), engine=engine)
      assert solver.engine() == engine
      snaps = solver.integrate(fixture[config + '/u0'], 0.0, dt, wl.FULL_STEPS, SAVE_EVERY).cpu().numpy().astype(np.float64)
      e = _errors(snaps, fixture[config + '/f32'], fixture[config + '/f64'])
      e['how'] = 'libddd1d, DDD1D_ENGINE_%s, B200' % engine.upper()
      results.setdefault(config, {})['gpu_' + engine] = e
      print(config, engine, e, flush=True)
      solver.close()
  merge(out_path, results, fixture)


def merge(out_path, results, fixture):
  data = {}
  if os.path.exists(out_path):
    with open(out_path) as f:
      data = json.load(f)
  for config, rows in results.items():
    data.setdefault(config, {}).update(rows)
  data['float32_oracle'] = {c: _errors(fixture[c + '/f32'], fixture[c + '/f32'], fixture[c + '/f64']) for c in CONFIGS}
  data['_about'] = ('relative L-inf per row (worst over rows 0/1/2047/4095 and the 10 snapshots of the 10 000-step horizon) '
                    'against the float64 oracle; float32_oracle = the reference arithmetic (float32 graph, float64 state)')
  with open(out_path, 'w') as f:
    json.dump(data, f, indent=1, sort_keys=True)


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--emulate', action='store_true')
  ap.add_argument('--gpu', action='store_true')
  ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'r02', 'tc_trajectory_error.json'))
  a = ap.parse_args()
  if a.emulate:
    run_emulation(a.out)
  if a.gpu:
    run_gpu(a.out)
