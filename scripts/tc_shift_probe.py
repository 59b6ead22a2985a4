"""Runs ddd1d_debug_tc_shift_probe (TMEM-resident A operand: tcgen05.cp / tcgen05.shift / A-in-TMEM MMA) and prints
what every step did.  Usage: gpurun -- python scripts/tc_shift_probe.py"""
import ctypes
import os
import numpy as np

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(root, 'data-driven-discretization-1d_b200', 'libddd1d_debug.so'))
out = np.zeros((10, 128, 16), np.uint32)
rc = lib.ddd1d_debug_tc_shift_probe(0, out.ctypes.data_as(ctypes.c_void_p))
print('rc', rc)
np.save(os.path.join(root, 'gpurun_out', 'tc_shift_probe.npy'), out) if os.path.isdir(os.path.join(root, 'gpurun_out')) else None


def decode(step, lanes):
  """raw pattern: each 32-bit column = two halfs (p | k << 8)."""
  for lane in lanes:
    cols = out[step, lane, :8]
    lo, hi = cols & 0xffff, cols >> 16
    print('  step %d lane %3d:' % (step, lane), ' '.join('p%d.k%d|p%d.k%d' % (l & 255, l >> 8, h & 255, h >> 8)
                                                         for l, h in zip(lo, hi)))


for step in range(5):
  print('step', step)
  decode(step, [0, 1, 31, 32, 127]) if step == 0 else None
  pos = (out[step, :, 0] & 0xff).astype(int)
  print('  position held by lane (col 0, low half):', pos.tolist())
  if step == 4:
    print('  columns 8..15, col 8 low half:', (out[step, :, 8] & 0xff).astype(int).tolist())

# MMA checks
p = np.arange(160)[:, None]
k = np.arange(16)[None, :]
A = (((p * 3 + k) % 7) - 3).astype(np.float64)          # [160][16]
n = np.arange(16)[:, None]
B = (((n + 2 * k) % 5) - 2).astype(np.float64)          # [16][16]
full = A @ B.T                                          # [160][16]: D row for the A row of position p
for step, name in ((5, 'plain MMA, A in TMEM'), (6, 'MMA.ashift'), (7, 'plain MMA after .ashift'),
                   (9, 'shift; MMA back to back')):
  D = out[step].view(np.float32).astype(np.float64)     # [128][16]
  # which A row does every lane's result correspond to?
  src = []
  for lane in range(128):
    hit = [q for q in range(160) if np.array_equal(full[q], D[lane])]
    src.append(hit[0] if len(hit) == 1 else (hit[:3] if hit else None))
  print(name, '-> A row per lane:', src)
a_after = out[8, :, :8]
halves = np.stack([a_after & 0xffff, a_after >> 16], -1).reshape(128, 16).astype(np.uint16).view(np.float16)
src = []
for lane in range(128):
  hit = [q for q in range(160) if np.array_equal(A[q], halves[lane].astype(np.float64))]
  src.append(hit[0] if len(hit) == 1 else (hit[:3] if hit else None))
print('A columns after .ashift + plain: A row per lane:', src)
