"""ddd1d_b200: B200-native explicit 1-D PDE integration with learned finite-difference
coefficients -- a drop-in for the time-integration path of
google/data-driven-discretization-1d (``pde_superresolution``).

Host-side mirror of the reference modules (same names): equations, polynomials,
training (create_hparams), model, integrate, weno, duckarray; the numerical work is
in libddd1d.so (csrc/, C ABI in include/ddd1d.h).  Importing this package loads no
CUDA context; the first solver construction does, and fails loudly without a GPU
or without the built library.
"""
from . import _lib
from . import duckarray
from . import polynomials
from . import equations
from . import training
from . import layers
from . import analysis
from . import checkpoint
from . import runtime
from . import model
from . import weno
from . import integrate
from . import distributed
from . import evaluation

__all__ = ['duckarray', 'polynomials', 'equations', 'training', 'checkpoint', 'runtime', 'model', 'weno',
           'integrate', 'distributed', 'evaluation']
