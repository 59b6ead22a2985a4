"""Does CUDA-core work overlap tcgen05 MMAs whose operands stream from shared memory?
csrc/ddd1d_tc.cuh: tc_overlap_kernel.  Prints clocks for each stream alone and together."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddd1d_b200 import _lib
lib = _lib.load_debug()
WORK = {1: ('FFMA chain', 400), 2: ('STS.128 + LDS.128', 2000), 3: ('packed fp16 splits', 800), 4: ('tcgen05.ld x16', 4000),
        5: ('SHFL', 2000), 6: ('LDG (L1 hits)', 2000), 7: ('LDS.128 only', 2000), 8: ('STS.128 only', 4000),
        9: ('mbarrier arrive+wait', 2000), 10: ('bar.sync 128', 4000),
        11: ('STS + fence.proxy.async', 2000), 12: ('tcgen05 fences', 4000), 13: ('plane store round', 2000)}
REPS = 100      # x 10 steps x 2 MMAs


def run(mode, iters, blocks):
  out = np.zeros(2 * blocks, np.int64)
  _lib.check(lib.ddd1d_debug_tc_overlap(0, mode, REPS, iters, blocks, _lib.host_ptr(out)))
  return out[0::2].mean(), out[1::2].mean()


for blocks in (1, 148):
  mma_alone, _ = run(1, 1, blocks)
  print('blocks %3d  MMAs alone %9.0f clk (%.1f per step)' % (blocks, mma_alone, mma_alone / (REPS * 10)))
  for w, (name, iters) in WORK.items():
    _, cuda_alone = run(w << 1, iters, blocks)
    mma_both, cuda_both = run((w << 1) | 1, iters, blocks)
    print('blocks %3d  %-20s alone %9.0f   with MMAs: cuda %9.0f  mma %9.0f   (serial sum %9.0f)' %
          (blocks, name, cuda_alone, cuda_both, mma_both, cuda_alone + mma_alone))
