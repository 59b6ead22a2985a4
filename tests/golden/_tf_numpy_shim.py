"""A NumPy-eager stand-in for the handful of TensorFlow-1 ops the reference's
inference graph uses, so that the reference's OWN graph-construction code
(model.predict_coefficients, model.baseline_space_derivatives, layers.pad_periodic,
polynomials.reconstruct, equations.*, weno.* on tf.Tensor inputs ...) can be
executed in the authoring container, where tensorflow<2 cannot be installed.

What is the reference's and what is ours:
  * everything above the primitive-op level (padding arithmetic, slicing, channel
    splitting, the order in which ops are combined, float32 casts of NumPy
    constants) is the reference's code, run unmodified from /root/reference;
  * the primitive ops below follow TensorFlow's documented semantics:
      - tf.layers.conv1d / tf.nn.conv1d: VALID cross-correlation, NWC input,
        kernel [width, in, out]  (out[b,x,o] = sum_{k,i} in[b,x+k,i] W[k,i,o] + b[o]);
      - tf.extract_image_patches: VALID patches, depth ordered (row, col, channel);
      - tf.einsum, concat, tile, stack, squeeze, reshape, reduce_*: as NumPy;
      - binary ops between a Tensor and a NumPy array / Python number convert the
        non-tensor operand to the tensor's dtype (tf.convert_to_tensor with
        preferred dtype), Tensor-Tensor dtype mismatches raise as in TF.
    Every op computes in float64 and rounds its output ONCE to the tensor dtype,
    so the golden values sit at the centre of whatever summation order the real
    float32 kernels (Eigen, cuDNN) would have used.

Used only by tests/golden/make_golden.py.  Never imported on the GPU box.
"""
import contextlib
import types

import numpy as np


class Dimension(object):
  def __init__(self, value):
    self.value = value


class TensorShape(object):
  def __init__(self, dims):
    self._dims = tuple(int(d) for d in dims)

  def __len__(self):
    return len(self._dims)

  def __getitem__(self, index):
    if isinstance(index, slice):
      return TensorShape(self._dims[index])
    return Dimension(self._dims[index])

  def __iter__(self):
    return iter(Dimension(d) for d in self._dims)

  def as_list(self):
    return list(self._dims)

  def concatenate(self, other):
    other = other.as_list() if isinstance(other, TensorShape) else list(other)
    return TensorShape(self._dims + tuple(other))


def _round(value64, dtype):
  return np.asarray(value64).astype(dtype)


class Tensor(object):
  """Eager float32/float64/bool/int tensor."""
  __array_ufunc__ = None  # make ndarray (op) Tensor defer to Tensor.__r<op>__

  def __init__(self, array, dtype=None):
    if isinstance(array, Tensor):
      array = array.a
    self.a = np.array(array, dtype=dtype)

  # -- structure --------------------------------------------------------------
  @property
  def shape(self):
    return TensorShape(self.a.shape)

  @property
  def dtype(self):
    return self.a.dtype

  def set_shape(self, shape):
    pass

  def __getitem__(self, index):
    return Tensor(self.a[index])

  def eval(self):
    return self.a

  # -- arithmetic -------------------------------------------------------------
  def _coerce(self, other):
    if isinstance(other, Tensor):
      if other.a.dtype != self.a.dtype:
        raise TypeError('dtype mismatch %s vs %s' % (self.a.dtype, other.a.dtype))
      return other.a
    return np.asarray(other, dtype=self.a.dtype)

  def _binary(self, other, fn, reverse=False):
    o = self._coerce(other)
    x, y = (o, self.a) if reverse else (self.a, o)
    if self.a.dtype.kind == 'f':
      return Tensor(_round(fn(x.astype(np.float64), y.astype(np.float64)), self.a.dtype))
    return Tensor(fn(x, y))

  def __add__(self, o): return self._binary(o, np.add)
  def __radd__(self, o): return self._binary(o, np.add, True)
  def __sub__(self, o): return self._binary(o, np.subtract)
  def __rsub__(self, o): return self._binary(o, np.subtract, True)
  def __mul__(self, o): return self._binary(o, np.multiply)
  def __rmul__(self, o): return self._binary(o, np.multiply, True)
  def __truediv__(self, o): return self._binary(o, np.true_divide)
  def __rtruediv__(self, o): return self._binary(o, np.true_divide, True)
  __div__ = __truediv__
  __rdiv__ = __rtruediv__

  def __pow__(self, o):
    return self._binary(o, np.power)

  def __neg__(self):
    return Tensor(-self.a)

  def _compare(self, other, fn):
    return Tensor(fn(self.a, self._coerce(other)))

  def __le__(self, o): return self._compare(o, np.less_equal)
  def __lt__(self, o): return self._compare(o, np.less)
  def __ge__(self, o): return self._compare(o, np.greater_equal)
  def __gt__(self, o): return self._compare(o, np.greater)


def _a(x):
  return x.a if isinstance(x, Tensor) else np.asarray(x)


def _unary(fn):
  def op(x, name=None):
    x = Tensor(x)
    return Tensor(_round(fn(x.a.astype(np.float64)), x.a.dtype))
  return op


class VariableStore(object):
  """Ordered (kernel, bias) pairs handed out to tf.layers.conv1d calls, in the
  creation order conv1d, conv1d_1, ... that tf.train.Saver would have used."""

  def __init__(self):
    self.pending = []
    self.named = {}


STORE = VariableStore()


def build():
  tf = types.ModuleType('tensorflow')
  tf._ddd1d_stub = True
  tf._ddd1d_numpy = True
  tf.Tensor = Tensor
  tf.float32, tf.float64, tf.int32 = np.float32, np.float64, np.int32
  tf.newaxis = None
  tf.AUTO_REUSE = None

  @contextlib.contextmanager
  def name_scope(name=None, default_name=None, values=None):
    yield name or default_name
  tf.name_scope = name_scope

  @contextlib.contextmanager
  def variable_scope(name, reuse=None):
    yield name
  tf.variable_scope = variable_scope

  def convert_to_tensor(value, dtype=None, name=None):
    if isinstance(value, Tensor):
      if dtype is not None and value.a.dtype != np.dtype(dtype):
        raise TypeError('dtype mismatch')
      return value
    return Tensor(value, dtype=dtype)
  tf.convert_to_tensor = convert_to_tensor
  tf.constant = lambda value, dtype=None, name=None: Tensor(value, dtype=dtype)
  tf.identity = lambda x, name=None: Tensor(x)

  tf.shape = lambda x: Tensor(np.array(_a(x).shape, dtype=np.int32))
  tf.tile = lambda x, multiples: Tensor(np.tile(_a(x), tuple(
      int(np.asarray(_a(m))) for m in (multiples if isinstance(multiples, (list, tuple)) else _a(multiples)))))
  tf.concat = lambda values, axis, name=None: Tensor(np.concatenate([_a(v) for v in values], axis=axis))
  tf.stack = lambda values, axis=0: Tensor(np.stack([_a(v) for v in values], axis=axis))
  tf.squeeze = lambda x, axis=None: Tensor(np.squeeze(_a(x), axis=axis))
  tf.reshape = lambda x, shape: Tensor(np.reshape(_a(x), tuple(int(s) for s in _a(shape))))
  tf.transpose = lambda x, perm=None: Tensor(np.transpose(_a(x), perm))

  def reduce(fn):
    def op(x, axis=None, keepdims=False, **unused):
      x = Tensor(x)
      return Tensor(_round(fn(x.a.astype(np.float64), axis=axis, keepdims=keepdims), x.a.dtype))
    return op
  tf.reduce_sum = reduce(np.sum)
  tf.reduce_mean = reduce(np.mean)

  tf.sin = _unary(np.sin)
  tf.tanh = _unary(np.tanh)
  tf.maximum = lambda x, y: Tensor(np.maximum(_a(x), _a(y)))
  tf.minimum = lambda x, y: Tensor(np.minimum(_a(x), _a(y)))
  tf.where = lambda c, x, y: Tensor(np.where(_a(c), _a(x), _a(y)))

  def einsum(equation, *inputs):
    dtype = _a(inputs[0]).dtype
    out = np.einsum(equation, *[_a(i).astype(np.float64) for i in inputs])
    return Tensor(_round(out, dtype))
  tf.einsum = einsum

  def _conv_valid(x, kernel):
    width = kernel.shape[0]
    n_out = x.shape[1] - width + 1
    acc = np.zeros((x.shape[0], n_out, kernel.shape[2]), dtype=np.float64)
    x64, k64 = x.astype(np.float64), kernel.astype(np.float64)
    for tap in range(width):
      acc += x64[:, tap:tap + n_out, :] @ k64[tap]
    return acc

  nn = types.SimpleNamespace()

  def nn_conv1d(value, filters, stride, padding, **unused):
    assert stride == 1 and padding == 'VALID'
    x, k = _a(value), _a(filters)
    assert x.dtype == k.dtype, (x.dtype, k.dtype)
    return Tensor(_round(_conv_valid(x, k), x.dtype))
  nn.conv1d = nn_conv1d
  nn.relu = lambda x: Tensor(np.maximum(_a(x), 0))
  nn.relu6 = lambda x: Tensor(np.minimum(np.maximum(_a(x), 0), 6).astype(_a(x).dtype))
  nn.softplus = _unary(lambda v: np.logaddexp(v, 0))
  nn.elu = _unary(lambda v: np.where(v > 0, v, np.expm1(np.minimum(v, 0))))
  tf.nn = nn

  layers = types.SimpleNamespace()

  def layers_conv1d(inputs, filters, kernel_size, strides=1, padding='valid',
                    dilation_rate=1, activation=None, **unused):
    assert strides == 1 and dilation_rate == 1 and padding == 'valid'
    x = _a(inputs)
    kernel, bias = STORE.pending.pop(0)
    assert kernel.shape == (kernel_size, x.shape[2], filters), (kernel.shape, kernel_size, x.shape, filters)
    assert kernel.dtype == x.dtype and bias.dtype == x.dtype
    out = Tensor(_round(_conv_valid(x, kernel) + bias.astype(np.float64), x.dtype))
    return activation(out) if activation is not None else out
  layers.conv1d = layers_conv1d
  tf.layers = layers

  def get_variable(name, shape, initializer=None):
    value = STORE.named[name]
    assert tuple(value.shape) == tuple(shape)
    return Tensor(value)
  tf.get_variable = get_variable
  tf.initializers = types.SimpleNamespace(zeros=lambda: None)

  def extract_image_patches(images, ksizes, strides, rates, padding):
    assert strides == [1, 1, 1, 1] and rates == [1, 1, 1, 1] and padding == 'VALID'
    x = _a(images)
    assert ksizes[0] == 1 and ksizes[2] == 1 and ksizes[3] == 1 and x.shape[2] == 1 and x.shape[3] == 1
    size = ksizes[1]
    n_out = x.shape[1] - size + 1
    out = np.stack([x[:, i:i + n_out, :, 0] for i in range(size)], axis=-1)   # [b, rows, cols, size]
    return Tensor(out)
  tf.extract_image_patches = extract_image_patches

  spectral = types.SimpleNamespace(
      rfft=lambda x: Tensor(np.fft.rfft(_a(x))), irfft=lambda x: Tensor(np.fft.irfft(_a(x))))
  tf.spectral = spectral

  class HParams(object):
    def __init__(self, **kwargs):
      self.__dict__.update(kwargs)

    def override_from_dict(self, values):
      for k, v in values.items():
        if k not in self.__dict__:
          raise ValueError('unknown hparam %s' % k)
        setattr(self, k, v)
      return self

    def values(self):
      return dict(self.__dict__)

  contrib = types.SimpleNamespace(training=types.SimpleNamespace(HParams=HParams))
  tf.contrib = contrib
  return tf
