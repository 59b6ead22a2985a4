"""Collector-buffer reuse of tcgen05.mma (ddd1d_debug_tc_reuse_rate): clocks per MMA for pairs sharing an operand.
Every configuration runs in its own process (an unsupported shape raises "illegal instruction" and kills the context).
Usage: gpurun -- python scripts/tc_reuse_rate.py"""
import ctypes
import os
import subprocess
import sys

import numpy as np

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ['plain pair, same B, different A', 'mma.ws pair, B kept in collector b0', 'same A, collector::a fill / lastuse',
         'same A, no qualifiers']
if len(sys.argv) == 3:
  n, mode = int(sys.argv[1]), int(sys.argv[2])
  lib = ctypes.CDLL(os.path.join(root, 'data-driven-discretization-1d_b200', 'libddd1d_debug.so'))
  out = np.zeros(1, np.int64)
  rc = lib.ddd1d_debug_tc_reuse_rate(0, n, mode, 200, out.ctypes.data_as(ctypes.c_void_p))
  print('N=%3d  %-40s rc %d  %6.1f clk per MMA' % (n, NAMES[mode], rc, out[0] / 400.0))
else:
  for n in (32, 64, 128):
    for mode in range(4):
      r = subprocess.run([sys.executable, __file__, str(n), str(mode)], capture_output=True, text=True)
      print((r.stdout.strip() or r.stderr.strip().splitlines()[-1])[:200])
