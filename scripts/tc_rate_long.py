"""MMA rate over long streams (does a sustained tcgen05 stream slow down?)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddd1d_b200 import _lib
lib = _lib.load_debug()
for blocks in (1, 148):
  for reps in (100, 1000, 10000, 60000):
    out = np.zeros(blocks, np.int64)
    _lib.check(lib.ddd1d_debug_tc_rate(0, 15, reps, blocks, _lib.host_ptr(out)))
    print('blocks %3d reps %6d  f16 64+32: %7.1f clk/step (max over blocks %7.1f)' % (blocks, reps, out.mean() / (reps * 20), out.max() / (reps * 20)))
