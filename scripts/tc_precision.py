"""Where does the tensor path lose bits?  Runs the tcgen05 probe (one 128x32x160 tile)
on inputs that isolate each error source and prints relative L-inf errors vs float64."""
import ctypes
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddd1d_b200 import _lib

lib = _lib.load_debug()
MASK = np.uint32(0xffffe000)

def rn_tf32(a):
  return ((a.view(np.uint32) + np.uint32(0x1000)) & MASK).view(np.float32)

def run(x, w, nout, lo_zero=False):
  packed = np.zeros((40, 2 * nout, 4), np.float32)
  hi = rn_tf32(np.ascontiguousarray(w))
  lo = np.zeros_like(hi) if lo_zero else rn_tf32((w - hi).astype(np.float32))
  for k in range(5):
    for ci in range(32):
      packed[k * 8 + ci // 4, :nout, ci % 4] = hi[k, ci, :]
      packed[k * 8 + ci // 4, nout:, ci % 4] = lo[k, ci, :]
  dx, dw = (torch.as_tensor(a).cuda().contiguous() for a in (x, packed))
  out = torch.zeros((128, nout), dtype=torch.float32, device='cuda')
  _lib.check(lib.ddd1d_debug_tc_probe(0, dx.data_ptr(), dw.data_ptr(), out.data_ptr(), nout,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
  torch.cuda.synchronize()
  return out.cpu().numpy()

def exact(x, w):
  out = np.zeros((128, w.shape[2]))
  for k in range(5):
    out += x[k:k + 128].astype(np.float64) @ w[k].astype(np.float64)
  return out

def f32_chain(x, w):
  """float32 FMA-free sequential accumulation in the kernel's (ci, k) order: what FP32 FFMA does."""
  out = np.zeros((128, w.shape[2]), np.float32)
  for ci in range(32):
    for k in range(5):
      out = (out + x[k:k + 128, ci:ci + 1] * w[k, ci][None, :]).astype(np.float32)
  return out

def rel(a, b):
  return float(np.abs(a - b).max() / np.abs(b).max())

for nout in (32, 16):
  for sign in ('mixed', 'positive'):
    rs = np.random.RandomState(7)
    x = rs.randn(132, 32).astype(np.float32)
    w = (rs.randn(5, 32, nout) / 8).astype(np.float32)
    if sign == 'positive':
      x, w = np.abs(x), np.abs(w)
    xr, wr = rn_tf32(x), rn_tf32(w)
    print('nout=%d %s' % (nout, sign))
    print('  full 3xTF32 vs exact            %.3e' % rel(run(x, w, nout), exact(x, w)))
    print('  pre-rounded inputs (acc only)   %.3e' % rel(run(xr, wr, nout, lo_zero=True), exact(xr, wr)))
    print('  x pre-rounded, w split          %.3e' % rel(run(xr, w, nout), exact(xr, w)))
    print('  x split, w pre-rounded          %.3e' % rel(run(x, wr, nout, lo_zero=True), exact(x, wr)))
    print('  float32 sequential chain (FFMA) %.3e' % rel(f32_chain(x, w), exact(x, w)))
