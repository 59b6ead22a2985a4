"""The synthetic BASELINE.json workloads: one definition shared by bench.py,
the long-horizon parity fixtures (tests/golden/make_long_horizon.py) and the gpu tests, so that the
trajectory the bench times is the trajectory the tests pin.  NumPy only; nothing here touches CUDA.
"""
import math

import numpy as np

WORKLOADS = {
    # name: (equation kind, variant, N, per-GPU batch, dt, mode)
    'c2': ('burgers', 'plain', 256, 4096, 1e-3, 'learned'),     # BASELINE config 2 (the headline)
    'c3': ('kdv', 'plain', 128, 4096, 2.5e-5, 'learned'),       # config 3
    'c4': ('ks', 'plain', 512, 4096, 1e-5, 'learned'),          # config 4 (32768 rows over 8 GPUs)
    'c5': ('burgers', 'godunov', 2048, 8192, 1e-4, 'weno'),     # config 5 (65536 rows over 8 GPUs)
    'c1b': ('burgers', 'plain', 64, 65536, 1e-2, 'fd'),         # config 1 batched
    'c2s': ('burgers', 'plain', 64, 16384, 1e-3, 'learned'),    # the paper's grid (notebooks/time-integration.ipynb:252)
}
NET_OUTPUTS = {'burgers': 9, 'kdv': 8, 'ks': 11}
FULL_STEPS = 10000               # the horizon BASELINE.json's configs name


def synthetic_weights(kind, last_layer_scale=1e-2, seed=0):
  """Seeded Glorot-uniform kernels (tf.layers.conv1d's default initialiser), zero biases, last
  layer scaled by 1e-2 so the scheme stays a small perturbation of the 7-point FD bias and
  integrates stably for 10 k steps: [(kernel [5, cin, cout], bias [cout])]."""
  rs = np.random.RandomState(seed)
  shapes = [(5, 1, 32), (5, 32, 32), (5, 32, NET_OUTPUTS[kind])]
  out = []
  for i, (k, cin, cout) in enumerate(shapes):
    limit = math.sqrt(6.0 / (k * cin + k * cout))
    w = rs.uniform(-limit, limit, size=(k, cin, cout))
    if i == len(shapes) - 1:
      w = w * last_layer_scale
    out.append((w.astype(np.float32), np.zeros(cout, np.float32)))
  return out


def initial_rows(batch, n, seed, workload=None):
  """Smooth O(1) periodic rows (three Fourier modes); zeros for the batched config 1, which is
  BurgersEquation.initial_value() (equations.py:256-257)."""
  if workload == 'c1b':
    return np.zeros((batch, n), np.float32)
  rs = np.random.RandomState(seed)
  x = 2 * np.pi * np.arange(n) / n
  rows = np.zeros((batch, n))
  for m in range(1, 4):
    rows += rs.randn(batch, 1) * np.sin(m * x + 2 * np.pi * rs.rand(batch, 1)) / m
  return (0.5 * rows).astype(np.float32)


# ---- the long-horizon parity cases (tests/golden/long_horizon.npz) ---------------------------------
HORIZON_PICKS = (0, 1, 2047, 4095)          # rows of the 4096-row batch the parity fixture holds
HORIZON_SEED = 1000                         # = bench.py's rank-0 seed


def horizon_rows(workload):
  """The 4096 initial rows of a long-horizon case.  Row 0 of the Burgers batch is the reference's own
  initial_value() = zeros (the tensor engine's activation bound starts at zero and must be re-calibrated
  as forcing grows the row); row 1 of the KS batch is a pure high-wavenumber mode that the u_xxxx term
  damps by many orders of magnitude (the bound must follow it down)."""
  kind, _, n, batch, _, _ = WORKLOADS[workload]
  rows = initial_rows(batch, n, HORIZON_SEED, workload)
  if kind == 'burgers':
    rows[0] = 0.0
  if kind == 'ks':
    rows[1] = np.sin(2 * np.pi * 40 * np.arange(n) / n).astype(np.float32)
  return rows


def decaying_rows(n=256):
  """Unforced Burgers rows that viscosity damps far below their initial amplitude within the horizon."""
  x = 2 * np.pi * np.arange(n) / n
  rows = np.stack([np.sin(8 * x), 0.7 * np.sin(12 * x + 0.3) + 0.2 * np.sin(9 * x),
                   1.5 * np.sin(6 * x), 0.05 * np.sin(3 * x) + np.sin(20 * x)])
  return rows.astype(np.float32)
