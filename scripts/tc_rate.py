"""tcgen05 MMA rate for small-N shapes from unswizzled K-major shared-memory operands (compile-time
patterns, no per-step issue overhead): clocks per (tap, ci-block) step.  csrc/ddd1d_tc.cuh: tc_rate_kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddd1d_b200 import _lib
lib = _lib.load_debug()
VARIANTS = ['tf32 N=16', 'tf32 N=32', 'tf32 N=64', 'tf32 N=128', 'tf32 N=32 of a 64-row B plane', 'tf32 N=32 single accumulator',
            'tf32 64+32 (hidden layer step)', 'tf32 32+16 (last layer step)', 'tf32 32+32',
            'bf16 N=32 (K=16)', 'bf16 N=64', 'bf16 N=96', 'bf16 96+64', 'bf16 N=128',
            'f16 64+32 separate D', 'f16 64+32 D overlap (production hidden)', 'f16 32+16 D overlap (production last)', 'f16 N=96 one D']
for blocks in (1, 148):
  for v, name in enumerate(VARIANTS):
    line = 'blocks %3d  %-40s' % (blocks, name)
    for reps in (100, 1):       # reps = 1: one 20-step job from a cold pipe (fill + drain latency included)
      out = np.zeros(blocks, np.int64)
      _lib.check(lib.ddd1d_debug_tc_rate(0, v, reps, blocks, _lib.host_ptr(out)))
      line += '  reps %3d: %7.1f clk/step' % (reps, out.mean() / (reps * 20))
    print(line)

# cost of tcgen05.commit inside an MMA stream (production commits once per tile = every 10 MMA pairs)
for v, name in ((15, 'no commit'), (18, 'commit every 10 steps'), (19, 'commit every 5 steps'), (20, 'commit every step')):
  out = np.zeros(1, np.int64)
  _lib.check(lib.ddd1d_debug_tc_rate(0, v, 100, 1, _lib.host_ptr(out)))
  print('f16 64+32 (hidden step), %-24s %7.1f clk/step' % (name, out.mean() / (100 * 20)))

# does the operand DATA change the rate?  (values near 1 | random bit patterns incl. NaN/Inf/subnormals | zeros)
for kind, name in ((0, 'values near 1'), (1, 'random bits'), (2, 'zeros')):
  out = np.zeros(148, np.int64)
  _lib.check(lib.ddd1d_debug_tc_rate(0, 15, 100 | (kind << 16), 148, _lib.host_ptr(out)))
  print('f16 64+32 (hidden step), data = %-16s %7.1f clk/step' % (name, out.mean() / (100 * 20)))
