"""ctypes binding of libddd1d.so (include/ddd1d.h).  There is no CPU fallback: if
the library is missing or a call fails this module raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libddd1d.so')
DEBUG_LIB_PATH = os.path.join(_HERE, 'libddd1d_debug.so')     # test-only laboratory kernels

OK, EINVAL, EUNSUPPORTED, ECUDA, ESTATE = 0, -1, -2, -3, -4
BURGERS, KDV, KS = 0, 1, 2
PLAIN, CONSERVATIVE, GODUNOV = 0, 1, 2
MODE_STENCIL, MODE_LEARNED, MODE_WENO = 0, 1, 2
ACTIVATIONS = {None: 0, 'none': 0, 'relu': 1, 'relu6': 2, 'tanh': 3, 'softplus': 4, 'elu': 5}
PROJ_NULLSPACE, PROJ_RAW, PROJ_RAW_UNBIASED = 0, 1, 2
PROJ_DERIVATIVES, PROJ_TIME_DERIVATIVE, PROJ_FLUX = 3, 4, 5
SCHEMES = {'rk3': 0, 'RK23': 0, 'bogacki_shampine': 0, 'midpoint': 1, 'euler': 2, 'rk4': 3}
REAL_F32, REAL_F64 = 0, 1
STATE_F32 = 0x100     # OR into a scheme: float32 carry between steps (model.integrate_ode semantics)
ENGINES = {'auto': 0, 'ffma': 1, 'tensor': 2, 'tensor_f16x2': 3, 'tensor_f16': 4}
ENGINE_NAMES = {v: k for k, v in ENGINES.items() if v}
WINDOW = 11         # DDD1D_WINDOW: offsets -5..+5


class Config(ctypes.Structure):
  _fields_ = [
      ('struct_bytes', ctypes.c_int), ('device', ctypes.c_int), ('equation', ctypes.c_int),
      ('variant', ctypes.c_int), ('mode', ctypes.c_int), ('num_points', ctypes.c_int),
      ('num_derivatives', ctypes.c_int), ('weno_real', ctypes.c_int),
      ('dx', ctypes.c_double), ('eta', ctypes.c_double), ('standard_deviation', ctypes.c_double),
      ('num_layers', ctypes.c_int), ('filter_size', ctypes.c_int), ('kernel_size', ctypes.c_int),
      ('activation', ctypes.c_int), ('net_outputs', ctypes.c_int), ('stencil_size', ctypes.c_int),
      ('projection', ctypes.c_int), ('engine', ctypes.c_int),
  ]


class LibraryMissing(RuntimeError):
  pass


class Ddd1dError(RuntimeError):
  """A libddd1d call failed (other than bad arguments, which raise ValueError like
  the reference)."""


_lib = None

_P = ctypes.c_void_p
_SIGNATURES = {
    'ddd1d_version': (ctypes.c_int, []),
    'ddd1d_last_error': (ctypes.c_char_p, [_P]),
    'ddd1d_create': (ctypes.c_int, [ctypes.POINTER(Config), ctypes.POINTER(_P)]),
    'ddd1d_destroy': (ctypes.c_int, [_P]),
    'ddd1d_set_stencils': (ctypes.c_int, [_P, _P]),
    'ddd1d_set_layer': (ctypes.c_int, [_P, ctypes.c_int, _P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'ddd1d_set_projection': (ctypes.c_int, [_P, _P, _P]),
    'ddd1d_set_forcing': (ctypes.c_int, [_P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_double]),
    'ddd1d_rhs': (ctypes.c_int, [_P, ctypes.c_double, _P, _P, ctypes.c_int, ctypes.c_int, _P]),
    'ddd1d_rhs_f64': (ctypes.c_int, [_P, ctypes.c_double, _P, _P, ctypes.c_int, ctypes.c_int, _P]),
    'ddd1d_coefficients': (ctypes.c_int, [_P, _P, _P, ctypes.c_int, _P]),
    'ddd1d_space_derivatives': (ctypes.c_int, [_P, _P, _P, ctypes.c_int, _P]),
    'ddd1d_integrate': (ctypes.c_int, [_P, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, _P, _P, _P, ctypes.c_int, ctypes.c_int, _P]),
    'ddd1d_integrate_adaptive': (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                ctypes.c_double, _P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_int, _P]),
    'ddd1d_rhs_host': (ctypes.c_int, [_P, ctypes.c_double, _P, _P, ctypes.c_int, ctypes.c_int]),
    'ddd1d_integrate_host': (ctypes.c_int, [_P, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, _P, _P, _P, ctypes.c_int,
                                            ctypes.c_int]),
    'ddd1d_weno_reconstruct': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _P, _P, _P, ctypes.c_int,
                                              ctypes.c_int, _P]),
    'ddd1d_launch_count': (ctypes.c_longlong, [_P]),
    'ddd1d_launch_shape': (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    'ddd1d_engine': (ctypes.c_int, [_P]),
}

# test-only entry points (include/ddd1d_debug.h), in their own library: load_debug()
_DEBUG_SIGNATURES = {
    'ddd1d_debug_tc_probe': (ctypes.c_int, [ctypes.c_int, _P, _P, _P, ctypes.c_int, _P]),
    'ddd1d_debug_tc_rate': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
    'ddd1d_debug_tc_overlap': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
    'ddd1d_debug_tc_shift_probe': (ctypes.c_int, [ctypes.c_int, _P]),
    'ddd1d_debug_tc_war_probe': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
    'ddd1d_debug_tc_ta_rate': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, _P]),
    'ddd1d_debug_tc_reuse_rate': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
}


def exported_symbols():
  """Every entry point include/ddd1d.h declares."""
  return sorted(_SIGNATURES)


def load():
  """dlopen the in-tree library (built by __graft_entry__.build()); loud on failure."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise LibraryMissing(
          '%s not found: run `python __graft_entry__.py` (nvcc, sm_100a) first; '
          'there is no CPU fallback' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
      fn = getattr(lib, name)
      fn.restype = restype
      fn.argtypes = argtypes
    if lib.ddd1d_version() != 2:
      raise Ddd1dError('libddd1d version %d, binding expects 2' % lib.ddd1d_version())
    _lib = lib
  return _lib


_debug_lib = None


def load_debug():
  """The laboratory kernels of the tensor engine (descriptor probe, MMA rate / overlap microbenchmarks):
  tests and scripts only, never loaded by the product path."""
  global _debug_lib
  if _debug_lib is None:
    if not os.path.exists(DEBUG_LIB_PATH):
      raise LibraryMissing('%s not found: run `python __graft_entry__.py` first' % DEBUG_LIB_PATH)
    lib = ctypes.CDLL(DEBUG_LIB_PATH)
    for name, (restype, argtypes) in _DEBUG_SIGNATURES.items():
      fn = getattr(lib, name)
      fn.restype = restype
      fn.argtypes = argtypes
    _debug_lib = lib
  return _debug_lib


def check(code, handle=None):
  if code == OK:
    return
  msg = load().ddd1d_last_error(handle).decode('utf-8', 'replace')
  if code == EINVAL:
    raise ValueError(msg)
  if code == EUNSUPPORTED:
    raise NotImplementedError(msg)
  raise Ddd1dError('libddd1d error %d: %s' % (code, msg))


def host_ptr(array):
  return array.ctypes.data_as(_P)


def to_window(stencil):
  """Place an s-point stencil in the 11-slot window (offsets -5..+5) using the
  reference's centred alignment: ceil((s-1)/2) points left of the output point
  (layers.py:76-79 via nn_conv1d_periodic(center=True))."""
  stencil = np.asarray(stencil, dtype=np.float64)
  s = stencil.shape[-1]
  left = -(-(s - 1) // 2)
  half = WINDOW // 2
  if left > half or (s - 1 - left) > half:
    raise NotImplementedError('stencil of %d points does not fit the %d-point window' % (s, WINDOW))
  out = np.zeros(stencil.shape[:-1] + (WINDOW,), dtype=np.float64)
  out[..., half - left:half - left + s] = stencil
  return out
