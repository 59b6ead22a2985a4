"""Clocks per tile-layer of the TMEM-resident-A pattern (ddd1d_debug_tc_ta_rate) next to the shared-memory-A
pattern of the production kernel (882 clk hidden, 792 clk last layer: profiles/r01/tc_mma_rate.txt).
Usage: gpurun -- python scripts/tc_ta_rate.py"""
import ctypes
import os
import numpy as np

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(root, 'data-driven-discretization-1d_b200', 'libddd1d_debug.so'))
for blocks in (1, 148):
  for nb in (32, 16):
    for issuers in (1, 2, 4):
      for flags, name in ((3, 'ashift + patches'), (2, 'ashift, no patches'), (0, 'no shift, no patches')):
        reps = 100
        out = np.zeros(blocks * 4, np.int64)
        rc = lib.ddd1d_debug_tc_ta_rate(0, nb, reps, issuers, flags, blocks, out.ctypes.data_as(ctypes.c_void_p))
        c = out.reshape(blocks, 4)[:, :issuers]
        print('blocks %3d  NB=%2d  issuers %d  %-22s rc %d  %7.1f clk per tile-layer per issuer, %7.1f aggregate'
              % (blocks, nb, issuers, name, rc, c.mean() / reps, c.max(axis=1).mean() / (reps * issuers)))
