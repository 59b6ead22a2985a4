"""Diagnostic (gpu): per-position coefficient error of the tensor engine vs the float64 oracle."""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import pde_oracle as O
from tests import gpu_helpers as G
from ddd1d_b200 import runtime

np.set_printoptions(linewidth=250, precision=1)
for n in (32, 64, 128):
  for engine in ('tensor', 'tensor_f16x2', 'tensor_f16'):
    for batch in (1, 9):
      eq = [G.product_equation('burgers', 'plain', n, seed=s) for s in range(batch)]
      oeq = G.oracle_equation('burgers', 'plain', n, seed=0)
      w = O.glorot_weights(oeq, O.NetSpec(), seed=0, last_layer_scale=0.1, bias_scale=0.1)
      solver = runtime.learned_solver(eq, G.product_hparams('burgers', 'plain', n), w, engine=engine)
      u = G.smooth_rows(batch, n, seed=7)
      c64 = O.predict_coefficients(u, oeq, O.NetSpec(), w, dtype=np.float64)
      got = solver.coefficients(u).cpu().numpy()
      err = np.abs(got - c64).max(axis=(2, 3)) / np.abs(c64).max()       # [batch, x]
      print('N=%d %s batch=%d: max %.2e' % (n, engine, batch, err.max()))
      if err.max() > 1e-4:
        for b in range(batch):
          print('  row %d' % b, ' '.join('%.0e' % e if e > 1e-5 else '.' for e in err[b]))
      solver.close()
