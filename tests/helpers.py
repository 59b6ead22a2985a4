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
