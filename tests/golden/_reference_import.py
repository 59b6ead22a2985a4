"""Import the reference's NumPy-capable modules WITHOUT TensorFlow.

Only usable in the authoring container (where /root/reference is mounted); never
imported by the gpu tests, smoke() or bench.py.  It installs a permissive stand-in
module named ``tensorflow`` (the reference only needs ``tf.Tensor`` for isinstance
dispatch in its NumPy branches, plus attribute access at def/import time) and
registers an empty ``pde_superresolution`` package so that submodules import
without running the real ``__init__`` (which pulls in apache_beam / xarray).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('DDD1D_REFERENCE_ROOT', '/root/reference')


class _Any(object):
  """Object that tolerates any attribute access / call (import-time only)."""

  def __init__(self, *args, **kwargs):
    # also lets ``class X(tf.train.SessionRunHook)`` succeed: an instance used as
    # a base makes type(base) == _Any the metaclass, i.e. _Any(name, bases, ns).
    pass

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    return _Any()

  def __call__(self, *args, **kwargs):
    return _Any()


def available():
  return os.path.isdir(os.path.join(REFERENCE_ROOT, 'pde_superresolution'))


def _install_tf_stub(numpy_tf=False):
  existing = sys.modules.get('tensorflow')
  if existing is not None and not getattr(existing, '_ddd1d_stub', False):
    return  # a real tensorflow is present; leave it alone
  if existing is not None:
    if numpy_tf and not getattr(existing, '_ddd1d_numpy', False):
      raise RuntimeError('the inert tensorflow stub is already installed; request '
                         'numpy_tf=True before the first reference import')
    return
  if numpy_tf:
    import _tf_numpy_shim
    tf = _tf_numpy_shim.build()
  else:
    tf = types.ModuleType('tensorflow')
    tf._ddd1d_stub = True

    class Tensor(object):  # nothing is ever an instance of this
      pass

    tf.Tensor = Tensor
    for name in ('contrib', 'nn', 'layers', 'initializers', 'spectral'):
      setattr(tf, name, _Any())
    tf.float32 = 'float32'
    tf.float64 = 'float64'
    tf.AUTO_REUSE = None
    tf.tanh = None
    tf.newaxis = None
  tf.__getattr__ = lambda name: _Any()   # anything else touched at def time
  sys.modules['tensorflow'] = tf
  # dotted imports made at module level by training.py:29-32
  for dotted in ('tensorflow.contrib', 'tensorflow.contrib.training',
                 'tensorflow.contrib.training.python',
                 'tensorflow.contrib.training.python.training',
                 'tensorflow.core', 'tensorflow.core.protobuf'):
    mod = types.ModuleType(dotted)
    mod.__path__ = []
    mod.__getattr__ = lambda name: _Any()
    sys.modules[dotted] = mod
  for dotted, attr in (('tensorflow.contrib.training.python.training', 'hparam_pb2'),
                       ('tensorflow.core.protobuf', 'config_pb2'),
                       ('tensorflow.core.protobuf', 'rewriter_config_pb2')):
    sub = types.ModuleType(dotted + '.' + attr)
    sub.__getattr__ = lambda name: _Any()
    sys.modules[dotted + '.' + attr] = sub
    setattr(sys.modules[dotted], attr, sub)
  try:
    import xarray  # noqa: F401  (integrate.py:29; only used to package results)
  except ImportError:
    xr = types.ModuleType('xarray')
    xr.__getattr__ = lambda name: _Any()
    sys.modules['xarray'] = xr


def load(*names, numpy_tf=False):
  """Return the requested reference submodules, e.g. load('polynomials').

  numpy_tf=True installs the NumPy-eager TensorFlow shim (_tf_numpy_shim.py) so the
  reference's TF-graph code runs; it must be requested on the first call.
  """
  if not available():
    raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
  _install_tf_stub(numpy_tf)
  if 'pde_superresolution' not in sys.modules:
    pkg = types.ModuleType('pde_superresolution')
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'pde_superresolution')]
    sys.modules['pde_superresolution'] = pkg
  mods = [importlib.import_module('pde_superresolution.' + n) for n in names]
  return mods[0] if len(mods) == 1 else tuple(mods)
