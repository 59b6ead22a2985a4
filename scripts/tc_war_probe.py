"""Write-after-read stress of the TMEM-resident A operand (ddd1d_debug_tc_war_probe).
Usage: gpurun -- python scripts/tc_war_probe.py"""
import ctypes
import os
import numpy as np

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(root, 'data-driven-discretization-1d_b200', 'libddd1d_debug.so'))
for chain in (1, 4, 16, 64):
  out = np.zeros(2 * 128 * 16 + 1, np.uint32)
  rc = lib.ddd1d_debug_tc_war_probe(0, chain, 200, out.ctypes.data_as(ctypes.c_void_p))
  print('chain', chain, 'rc', rc, 'mismatches over 200 rounds:', int(out[-1]),
        'D0[0,:4]', out[:4].view(np.float32), 'D1[0,:4]', out[2048:2052].view(np.float32))
