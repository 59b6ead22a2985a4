"""TensorFlow-1 `tf.train.Saver` checkpoints and `hparams.pbtxt` without TensorFlow.

The reference restores its conv weights with `tf.train.Saver().restore(sess,
<dir>/model.ckpt)` (integrate.py:66-68, training.py:510-524) and its hyper-parameters
from `<dir>/hparams.pbtxt`, the text form of `HParams.to_proto()`
(training.py:590-592, 639-647).  TensorFlow is a third-party dependency of the
reference that is absent here (`tensorflow<2`, setup.py:26), so this module restates
the two published on-disk formats:

* checkpoint V2 "tensor bundle" (tensorflow/core/util/tensor_bundle): `<prefix>.index`
  is a LevelDB-format table (tensorflow/core/lib/io/table: prefix-compressed blocks,
  restart array, 1-byte compression tag (none | snappy, both read) + masked CRC-32C trailer, 48-byte footer with
  magic 0xdb4775248b80fb57) mapping "" -> BundleHeaderProto and every variable name ->
  BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c};
  `<prefix>.data-00000-of-00001` holds the raw little-endian tensor bytes;
* HParamDef text proto: `hparam { key: "..." value { int64_value | float_value |
  bytes_value | bool_value | int64_list{value..} | float_list | bytes_list | bool_list } }`.

PARITY UNPINNED: the reference ships no checkpoint file and TensorFlow cannot run here,
so the table reader is pinned only against the writer below (same format description), the
format's CRC-32C (RFC 3720 vectors) and magic; the protobuf wire encoding of the entries and
the hparams text format ARE cross-checked against the protobuf runtime (tests/test_checkpoint.py).
The first real `model.ckpt` should be read with `verify=True`.

Variables of the coefficient model live under scope `predict_coefficients/`
(model.py:442) with tf.layers naming: `conv1d/kernel [k, in, out]`, `conv1d/bias`,
`conv1d_1/...`, ... in creation order; `num_layers=0` models hold one `coefficients`
variable (model.py:497-500).  Optimiser slots (`.../Adam`, `beta1_power`, `global_step`)
are ignored.
"""
import os
import re
import struct

import numpy as np

_TABLE_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}     # tensorflow DataType enum
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------------------------
# CRC-32C (Castagnoli), masked as LevelDB / TensorFlow store it
# ---------------------------------------------------------------------------------
def _crc_table():
  table = np.zeros(256, dtype=np.uint32)
  for i in range(256):
    c = i
    for _ in range(8):
      c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    table[i] = c
  return table


_CRC_TABLE = _crc_table()


def crc32c(data, crc=0):
  crc ^= 0xFFFFFFFF
  table = _CRC_TABLE
  for b in bytes(data):
    crc = int(table[(crc ^ b) & 0xFF]) ^ (crc >> 8)
  return crc ^ 0xFFFFFFFF


def mask_crc(crc):
  return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------
# protobuf wire format (just what the two bundle messages need)
# ---------------------------------------------------------------------------------
def _read_varint(buf, pos):
  result, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def _write_varint(value):
  out = bytearray()
  while True:
    b = value & 0x7F
    value >>= 7
    if value:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _parse_message(buf):
  """-> list of (field_number, wire_type, value); nested messages stay bytes."""
  fields, pos = [], 0
  while pos < len(buf):
    tag, pos = _read_varint(buf, pos)
    field, wire = tag >> 3, tag & 7
    if wire == 0:
      value, pos = _read_varint(buf, pos)
    elif wire == 1:
      value = struct.unpack_from('<Q', buf, pos)[0]
      pos += 8
    elif wire == 2:
      n, pos = _read_varint(buf, pos)
      value = bytes(buf[pos:pos + n])
      pos += n
    elif wire == 5:
      value = struct.unpack_from('<I', buf, pos)[0]
      pos += 4
    else:
      raise ValueError('unsupported protobuf wire type %d' % wire)
    fields.append((field, wire, value))
  return fields


def _field(field, wire, payload):
  tag = _write_varint((field << 3) | wire)
  if wire == 0:
    return tag + _write_varint(payload)
  if wire == 2:
    return tag + _write_varint(len(payload)) + payload
  if wire == 5:
    return tag + struct.pack('<I', payload)
  raise ValueError(wire)


def _parse_entry(buf):
  """BundleEntryProto -> dict(dtype, shape, shard_id, offset, size, crc32c)."""
  entry = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
  for field, _, value in _parse_message(buf):
    if field == 1:
      entry['dtype'] = value
    elif field == 2:            # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
      dims = []
      for f2, _, v2 in _parse_message(value):
        if f2 == 2:
          size = 0
          for f3, _, v3 in _parse_message(v2):
            if f3 == 1:
              size = v3
          dims.append(size)
      entry['shape'] = tuple(dims)
    elif field == 3:
      entry['shard_id'] = value
    elif field == 4:
      entry['offset'] = value
    elif field == 5:
      entry['size'] = value
    elif field == 6:
      entry['crc32c'] = value
    elif field == 7:
      entry['sliced'] = True
  return entry


def _encode_entry(array, offset, crc):
  shape = b''.join(_field(2, 2, _field(1, 0, int(d))) for d in array.shape)
  return (_field(1, 0, _DTYPE_CODES[array.dtype]) + _field(2, 2, shape) + _field(4, 0, offset) +
          _field(5, 0, array.nbytes) + _field(6, 5, crc))


# ---------------------------------------------------------------------------------
# LevelDB-format table
# ---------------------------------------------------------------------------------
def snappy_decompress(data):
  """Raw Snappy block (format_description.txt of google/snappy): varint uncompressed length, then
  literals (tag & 3 == 0) and back-references with 1-, 2- or 4-byte offsets.  tf.train.Saver writes its
  index uncompressed (tensor_bundle.cc sets kNoCompression), but the table format allows type-1 blocks."""
  total, pos = _read_varint(data, 0)
  out = bytearray()
  while pos < len(data):
    tag = data[pos]
    pos += 1
    kind = tag & 3
    if kind == 0:
      length = (tag >> 2) + 1
      if length > 60:                       # 61..64: the length follows in 1..4 little-endian bytes
        extra = length - 60
        length = int.from_bytes(data[pos:pos + extra], 'little') + 1
        pos += extra
      out += data[pos:pos + length]
      pos += length
      continue
    if kind == 1:
      length = ((tag >> 2) & 7) + 4
      offset = ((tag >> 5) << 8) | data[pos]
      pos += 1
    elif kind == 2:
      length = (tag >> 2) + 1
      offset = int.from_bytes(data[pos:pos + 2], 'little')
      pos += 2
    else:
      length = (tag >> 2) + 1
      offset = int.from_bytes(data[pos:pos + 4], 'little')
      pos += 4
    if offset == 0 or offset > len(out):
      raise ValueError('corrupt snappy block')
    for _ in range(length):                 # byte by byte: references may overlap their own output
      out.append(out[-offset])
  if len(out) != total:
    raise ValueError('corrupt snappy block: {} bytes, header says {}'.format(len(out), total))
  return bytes(out)


def _read_block(data, offset, size, verify):
  block = data[offset:offset + size]
  kind = data[offset + size]
  if verify:
    stored = struct.unpack_from('<I', data, offset + size + 1)[0]
    actual = mask_crc(crc32c(data[offset:offset + size + 1]))
    if stored != actual:
      raise ValueError('checkpoint index block at %d fails its CRC-32C' % offset)
  if kind == 1:
    return snappy_decompress(block)
  if kind != 0:
    raise ValueError('unknown block compression tag %d' % kind)
  return block


def _block_entries(block):
  num_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
  limit = len(block) - 4 - 4 * num_restarts
  pos, key, out = 0, b'', []
  while pos < limit:
    shared, pos = _read_varint(block, pos)
    unshared, pos = _read_varint(block, pos)
    value_len, pos = _read_varint(block, pos)
    key = key[:shared] + bytes(block[pos:pos + unshared])
    pos += unshared
    out.append((key, bytes(block[pos:pos + value_len])))
    pos += value_len
  return out


def read_index(path, verify=True):
  """`<prefix>.index` -> (header fields, {name: entry dict})."""
  with open(path, 'rb') as f:
    data = f.read()
  if len(data) < 48 or struct.unpack_from('<Q', data, len(data) - 8)[0] != _TABLE_MAGIC:
    raise ValueError('%s is not a TensorFlow checkpoint index (bad table magic)' % path)
  footer = data[-48:]
  pos = 0
  _, pos = _read_varint(footer, pos)          # metaindex handle
  _, pos = _read_varint(footer, pos)
  index_offset, pos = _read_varint(footer, pos)
  index_size, pos = _read_varint(footer, pos)
  entries, header = {}, None
  for _, handle in _block_entries(_read_block(data, index_offset, index_size, verify)):
    block_offset, p = _read_varint(handle, 0)
    block_size, p = _read_varint(handle, p)
    for key, value in _block_entries(_read_block(data, block_offset, block_size, verify)):
      if key == b'':
        header = _parse_message(value)
      else:
        entries[key.decode('utf-8')] = _parse_entry(value)
  return header, entries


def read_checkpoint(prefix, verify=True):
  """Every variable of a V2 checkpoint: {name: ndarray}.  `prefix` is what Saver.save
  was given, e.g. `<dir>/model.ckpt`."""
  header, entries = read_index(prefix + '.index', verify)
  num_shards = 1
  for field, _, value in header or []:
    if field == 1:
      num_shards = value
    if field == 2 and value != 0:
      raise NotImplementedError('big-endian checkpoints are not supported')
  shards = {}
  out = {}
  for name, e in sorted(entries.items()):
    if e['sliced']:
      raise NotImplementedError('partitioned variable %r is not supported' % name)
    if e['dtype'] not in _DTYPES:
      continue                      # strings etc.: nothing the model needs
    sid = e['shard_id']
    if sid not in shards:
      with open('%s.data-%05d-of-%05d' % (prefix, sid, num_shards), 'rb') as f:
        shards[sid] = f.read()
    raw = shards[sid][e['offset']:e['offset'] + e['size']]
    if len(raw) != e['size']:
      raise ValueError('checkpoint data shard is truncated at %r' % name)
    if verify and e['crc32c'] is not None and mask_crc(crc32c(raw)) != e['crc32c']:
      raise ValueError('tensor %r fails its CRC-32C' % name)
    out[name] = np.frombuffer(raw, dtype=np.dtype(_DTYPES[e['dtype']]).newbyteorder('<')).reshape(e['shape']).copy()
  return out


def _build_block(items, restart_interval=16):
  buf, restarts, last = bytearray(), [], b''
  for i, (key, value) in enumerate(items):
    shared = 0
    if i % restart_interval == 0:
      restarts.append(len(buf))
    else:
      while shared < min(len(key), len(last)) and key[shared] == last[shared]:
        shared += 1
    buf += _write_varint(shared) + _write_varint(len(key) - shared) + _write_varint(len(value))
    buf += key[shared:] + value
    last = key
  if not restarts:
    restarts = [0]
  for r in restarts:
    buf += struct.pack('<I', r)
  buf += struct.pack('<I', len(restarts))
  return bytes(buf)


def write_checkpoint(prefix, tensors, entries_per_block=None):
  """Write {name: ndarray} as a single-shard V2 checkpoint (uncompressed index).  Used
  by the tests and to hand weights trained elsewhere to code that expects `model.ckpt`.
  `entries_per_block` splits the index into several data blocks (TensorFlow starts a new
  block every 256 KiB of entries; the tests use it to exercise the two-level lookup)."""
  data, items = bytearray(), []
  header = _field(1, 0, 1) + _field(3, 2, _field(1, 0, 1))        # num_shards = 1, version.producer = 1
  items.append((b'', header))
  for name in sorted(tensors):
    array = np.asarray(tensors[name], order='C')         # (ascontiguousarray would turn scalars into shape (1,))
    if array.dtype not in _DTYPE_CODES:
      raise ValueError('unsupported dtype %s for %r' % (array.dtype, name))
    raw = array.astype(array.dtype.newbyteorder('<'), copy=False).tobytes()
    items.append((name.encode('utf-8'), _encode_entry(array, len(data), mask_crc(crc32c(raw)))))
    data += raw
  with open(prefix + '.data-00000-of-00001', 'wb') as f:
    f.write(bytes(data))

  out = bytearray()

  def emit(block):
    offset = len(out)
    out.extend(block)
    out.append(0)                                                  # no compression
    out.extend(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
    return _write_varint(offset) + _write_varint(len(block))

  step = entries_per_block or len(items)
  index_items = []
  for start in range(0, len(items), step):
    chunk = items[start:start + step]
    # index key: any string >= the block's last key and < the next block's first key
    index_items.append((chunk[-1][0] + b'\x00', emit(_build_block(chunk))))
  meta_handle = emit(_build_block([]))
  index_handle = emit(_build_block(index_items, restart_interval=1))
  footer = meta_handle + index_handle
  footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', _TABLE_MAGIC)
  out.extend(footer)
  with open(prefix + '.index', 'wb') as f:
    f.write(bytes(out))


def checkpoint_dir_to_path(checkpoint_dir):
  """training.py:523-524."""
  return os.path.join(checkpoint_dir, 'model.ckpt')


# Only predict_coefficients opens a variable scope (model.py:442).  The direct model targets
# ('space_derivatives' / 'time_derivative' / 'flux') build their conv stack through _multilayer_conv1d
# (model.py:551-568) with no scope, so their Saver variables are top-level `conv1d/kernel`, `conv1d_1/kernel`, ...
# Optimiser slots (`.../kernel/Adam`, `.../kernel/Adam_1`) do not end in kernel|bias and never match.
_LAYER = re.compile(r'^(?:.*/)?(?:predict_coefficients/)?conv1d(?:_(\d+))?/(kernel|bias)$')


def conv_weights(variables):
  """{name: ndarray} -> [(kernel[k, in, out], bias[out]), ...] in layer order, or
  [coefficients] for a `num_layers=0` model (model.py:497-500)."""
  layers = {}
  scoped = any(re.match(r'^(?:.*/)?predict_coefficients/', name) for name in variables)
  for name, value in variables.items():
    m = _LAYER.match(name)
    if m and (('predict_coefficients/' in name) == scoped):     # never mix a scoped stack with a stray unscoped one
      layers.setdefault(int(m.group(1) or 0), {})[m.group(2)] = np.asarray(value, np.float32)
  if not layers:
    for name, value in variables.items():
      if re.match(r'^(?:.*/)?predict_coefficients/coefficients$', name):
        return [np.asarray(value, np.float32)]
    raise ValueError('no [predict_coefficients/]conv1d* variables in the checkpoint (found: %s)'
                     % ', '.join(sorted(variables)[:8]))
  out = []
  for i in range(len(layers)):
    if i not in layers or set(layers[i]) != {'kernel', 'bias'}:
      raise ValueError('checkpoint is missing kernel/bias of conv layer %d' % i)
    out.append((layers[i]['kernel'], layers[i]['bias']))
  return out


def load_conv_weights(checkpoint_dir, verify=True):
  """The conv stack of the model saved in `checkpoint_dir` (a directory holding
  `model.ckpt.*`, or a checkpoint prefix)."""
  prefix = checkpoint_dir
  if os.path.isdir(checkpoint_dir):
    prefix = checkpoint_dir_to_path(checkpoint_dir)
    if not os.path.exists(prefix + '.index'):
      # fall back to the newest `model.ckpt-<step>` written by MonitoredTrainingSession
      found = sorted((int(m.group(1)), m.group(0)[:-6]) for m in
                     (re.match(r'^model\.ckpt-(\d+)\.index$', f) for f in os.listdir(checkpoint_dir)) if m)
      if not found:
        raise FileNotFoundError('no model.ckpt*.index in %s' % checkpoint_dir)
      prefix = os.path.join(checkpoint_dir, found[-1][1])
  return conv_weights(read_checkpoint(prefix, verify))


def save_conv_weights(checkpoint_dir, weights, model_target='coefficients'):
  """Inverse of load_conv_weights: `<dir>/model.ckpt.*` with the reference's variable names -- under
  `predict_coefficients/` for model_target='coefficients' (model.py:442), top level for the direct targets
  (model.py:551-568); a `num_layers=0` model is the single vector `predict_coefficients/coefficients`
  (model.py:497-500)."""
  os.makedirs(checkpoint_dir, exist_ok=True)
  tensors = {}
  prefix = 'predict_coefficients/' if model_target == 'coefficients' else ''
  if len(weights) == 1 and not isinstance(weights[0], (tuple, list)):
    if model_target != 'coefficients':
      raise ValueError('a num_layers=0 model has model_target="coefficients"')
    tensors['predict_coefficients/coefficients'] = np.asarray(weights[0], np.float32)
  else:
    for i, (kernel, bias) in enumerate(weights):
      scope = prefix + 'conv1d' + ('_%d' % i if i else '')
      tensors[scope + '/kernel'] = np.asarray(kernel, np.float32)
      tensors[scope + '/bias'] = np.asarray(bias, np.float32)
  write_checkpoint(checkpoint_dir_to_path(checkpoint_dir), tensors)


# ---------------------------------------------------------------------------------
# hparams.pbtxt
# ---------------------------------------------------------------------------------
_TOKEN = re.compile(r'\s*(?:(#[^\n]*)|([A-Za-z_][A-Za-z0-9_]*)|("(?:[^"\\]|\\.)*"|\'(?:[^\'\\]|\\.)*\')|'
                    r'([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf|nan)f?)|([{}:<>\[\],;]))')


def _tokens(text):
  pos, out = 0, []
  text = text.rstrip()
  while pos < len(text):
    m = _TOKEN.match(text, pos)
    if not m:
      raise ValueError('cannot parse hparams text near %r' % text[pos:pos + 30])
    pos = m.end()
    if m.group(1):
      continue
    out.append(m.group(2) or m.group(3) or m.group(4) or m.group(5))
  return out


def _unescape(token):
  body = token[1:-1]
  return bytes(body, 'utf-8').decode('unicode_escape').encode('latin-1').decode('utf-8')


def _parse_text_message(tokens, pos, closer=None):
  """-> (list of (name, value), next position); value is a scalar token or a nested list."""
  fields = []
  while pos < len(tokens):
    tok = tokens[pos]
    if closer and tok == closer:
      return fields, pos + 1
    name = tok
    pos += 1
    if tokens[pos] == ':':
      pos += 1
    if tokens[pos] in ('{', '<'):
      value, pos = _parse_text_message(tokens, pos + 1, '}' if tokens[pos] == '{' else '>')
    else:
      value = tokens[pos]
      pos += 1
    fields.append((name, value))
    if pos < len(tokens) and tokens[pos] in (',', ';'):
      pos += 1
  if closer:
    raise ValueError('unbalanced braces in hparams text')
  return fields, pos


def _scalar(kind, token):
  if kind.startswith('int64'):
    return int(token)
  if kind.startswith('float'):
    return float(token.rstrip('f')) if token not in ('inf', 'nan', '-inf') else float(token)
  if kind.startswith('bool'):
    return token in ('true', 'True', '1')
  if kind.startswith('bytes'):
    return _unescape(token)
  raise ValueError('unknown hparam value kind %r' % kind)


def parse_hparams_pbtxt(text):
  """HParamDef text proto -> {name: value} (lists for the *_list kinds)."""
  fields, _ = _parse_text_message(_tokens(text), 0)
  out = {}
  for name, value in fields:
    if name != 'hparam':
      continue
    entry = dict(value)
    key = _unescape(entry['key'])
    (kind, payload), = entry['value'] or [('bytes_value', '""')]
    if kind.endswith('_list'):
      out[key] = [_scalar(kind, tok) for n, tok in payload if n == 'value']
    else:
      out[key] = _scalar(kind, payload)
  return out


def _c_escape(data):
  """protobuf text_format's CEscape (as_utf8=False): printable ASCII as is, the usual C escapes,
  everything else as three-digit octal."""
  table = {9: '\\t', 10: '\\n', 13: '\\r', 34: '\\"', 39: "\\'", 92: '\\\\'}
  return ''.join(table.get(b) or (chr(b) if 32 <= b < 127 else '\\%03o' % b) for b in data)


def format_hparams_pbtxt(values):
  """{name: value} -> the text `str(HParams.to_proto())` produces (map entries sorted by key)."""
  def kind_of(v):
    if isinstance(v, bool):
      return 'bool'
    if isinstance(v, int):
      return 'int64'
    if isinstance(v, float):
      return 'float'
    if isinstance(v, str):
      return 'bytes'
    raise ValueError('unsupported hparam value %r' % (v,))

  def fmt(kind, v):
    if kind == 'bool':
      return 'true' if v else 'false'
    if kind == 'bytes':
      return '"%s"' % _c_escape(v.encode('utf-8'))
    if kind == 'float':
      # the float32 value's float64 repr: what the protobuf runtime installed here prints for a float field (pinned by
      # tests/test_checkpoint.py against google.protobuf.text_format); newer runtimes print the shortest float32 form,
      # which parse_hparams reads just the same
      return repr(float(np.float32(v))) if np.isfinite(v) else str(v)
    return str(v)

  lines = []
  for key in sorted(values):
    v = values[key]
    lines += ['hparam {', '  key: "%s"' % key, '  value {']
    if isinstance(v, (list, tuple)):
      kind = kind_of(v[0]) if v else 'float'
      lines.append('    %s_list {' % kind)
      lines += ['      value: %s' % fmt(kind, x) for x in v]
      lines.append('    }')
    else:
      kind = kind_of(v)
      lines.append('    %s_value: %s' % (kind, fmt(kind, v)))
    lines += ['  }', '}']
  return '\n'.join(lines) + '\n'
