"""Shared helpers for the test-suite (oracle side)."""
import json

import numpy as np

from oracle import pde_oracle as O

KINDS = ('burgers', 'kdv', 'ks')
VARIANTS = ('plain', 'conservative', 'godunov')


def weights_from(npz, prefix):
  out, i = [], 0
  while '%s/kernel%d' % (prefix, i) in npz.files:
    out.append((npz['%s/kernel%d' % (prefix, i)], npz['%s/bias%d' % (prefix, i)]))
    i += 1
  return out


def net_from_json(text):
  return O.NetSpec(**json.loads(str(text)))


def rel_err(actual, expected):
  actual = np.asarray(actual, dtype=np.float64)
  expected = np.asarray(expected, dtype=np.float64)
  scale = np.max(np.abs(expected))
  return np.max(np.abs(actual - expected)) / (scale if scale > 0 else 1.0)


def assert_f32_faithful(got, want32, want64, factor=3.0, floor=2e-6, what=''):
  """`got` (a float32 GPU result) must be as close to the exact answer (`want64`, the oracle
  evaluated in float64) as the reference's own float32 graph (`want32`) is, up to `factor`.
  Two correct float32 evaluations with different rounding order differ from each other by about
  the float32 graph's own error, so this -- not a fixed distance to `want32` -- is the meaningful
  bound wherever cancellation amplifies rounding (stencils scaled by 1/dx^n, flux differences)."""
  e_ref = rel_err(want32, want64)
  e_got = rel_err(got, want64)
  assert e_got <= max(factor * e_ref, floor), '%s: error %.3e vs float64, float32 reference graph has %.3e' % (
      what, e_got, e_ref)
  return e_got, e_ref
