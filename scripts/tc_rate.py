"""tcgen05 MMA rate for small-N shapes from unswizzled K-major shared-memory operands (compile-time
patterns, no per-step issue overhead): clocks per (tap, ci-block) step.  csrc/ddd1d_tc.cuh: tc_rate_kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddd1d_b200 import _lib
lib = _lib.load()
VARIANTS = ['tf32 N=16', 'tf32 N=32', 'tf32 N=64', 'tf32 N=128', 'tf32 N=32 of a 64-row B plane', 'tf32 N=32 single accumulator',
            'tf32 64+32 (hidden layer step)', 'tf32 32+16 (last layer step)', 'tf32 32+32',
            'bf16 N=32 (K=16)', 'bf16 N=64', 'bf16 N=96', 'bf16 96+64', 'bf16 N=128']
for blocks in (1, 148):
  for v, name in enumerate(VARIANTS):
    out = np.zeros(blocks, np.int64)
    reps = 100
    _lib.check(lib.ddd1d_debug_tc_rate(0, v, reps, blocks, _lib.host_ptr(out)))
    print('blocks %3d  %-34s %7.1f clk/step' % (blocks, name, out.mean() / (reps * 20)))
