"""TensorFlow-1 checkpoint bundle + hparams.pbtxt ingestion.

Parity unpinned (the reference ships no checkpoint and TensorFlow cannot run here):
the reader is checked against the writer that restates the same published format,
against the format's own CRCs and magic, and against a hand-assembled index block
that uses key prefix compression the writer never emits for the first key.
"""
import os
import struct

import numpy as np
import pytest

import ddd1d_b200 as ddd
from ddd1d_b200 import checkpoint as C


def _weights(seed=0, outputs=9):
  rs = np.random.RandomState(seed)
  shapes = [(5, 1, 32), (5, 32, 32), (5, 32, outputs)]
  return [(rs.randn(*s).astype(np.float32), rs.randn(s[-1]).astype(np.float32)) for s in shapes]


def test_crc32c_known_answers():
  # RFC 3720 B.4 test vectors
  assert C.crc32c(b'\x00' * 32) == 0x8A9136AA
  assert C.crc32c(b'\xff' * 32) == 0x62A8AB43
  assert C.crc32c(bytes(range(32))) == 0x46DD794E
  assert C.crc32c(b'123456789') == 0xE3069283
  # leveldb's crc32c_test: Mask is not the identity and is invertible in principle
  assert C.mask_crc(C.crc32c(b'foo')) != C.crc32c(b'foo')


def test_checkpoint_round_trip(tmp_path):
  w = _weights()
  d = str(tmp_path / 'ckpt')
  C.save_conv_weights(d, w)
  assert sorted(os.listdir(d)) == ['model.ckpt.data-00000-of-00001', 'model.ckpt.index']
  got = C.load_conv_weights(d)
  assert len(got) == 3
  for (k0, b0), (k1, b1) in zip(w, got):
    np.testing.assert_array_equal(k0, k1)
    np.testing.assert_array_equal(b0, b1)
  # the table footer carries the LevelDB magic
  raw = open(os.path.join(d, 'model.ckpt.index'), 'rb').read()
  assert struct.unpack('<Q', raw[-8:])[0] == 0xdb4775248b80fb57


def test_checkpoint_ignores_optimizer_slots_and_orders_layers(tmp_path):
  w = _weights(1)
  tensors = {'global_step': np.asarray(123, np.int64), 'beta1_power': np.asarray(0.5, np.float32)}
  for i, (k, b) in enumerate(w):
    scope = 'predict_coefficients/conv1d' + ('_%d' % i if i else '')
    tensors[scope + '/kernel'] = k
    tensors[scope + '/bias'] = b
    tensors[scope + '/kernel/Adam'] = np.zeros_like(k)
    tensors[scope + '/kernel/Adam_1'] = np.zeros_like(k)
    tensors[scope + '/bias/Adam'] = np.zeros_like(b)
  prefix = str(tmp_path / 'model.ckpt-20')
  C.write_checkpoint(prefix, tensors)
  everything = C.read_checkpoint(prefix)
  assert everything['global_step'] == 123 and everything['global_step'].dtype == np.int64
  assert len(everything) == len(tensors)
  got = C.load_conv_weights(str(tmp_path))          # directory without model.ckpt: newest model.ckpt-<step>
  for (k0, b0), (k1, b1) in zip(w, got):
    np.testing.assert_array_equal(k0, k1)
    np.testing.assert_array_equal(b0, b1)


def test_checkpoint_with_several_index_blocks(tmp_path):
  w = _weights(3, outputs=11)
  tensors = {}
  for i, (k, b) in enumerate(w):
    scope = 'predict_coefficients/conv1d' + ('_%d' % i if i else '')
    tensors[scope + '/kernel'], tensors[scope + '/bias'] = k, b
  prefix = str(tmp_path / 'model.ckpt')
  C.write_checkpoint(prefix, tensors, entries_per_block=2)        # header + 6 tensors -> 4 data blocks
  got = C.conv_weights(C.read_checkpoint(prefix))
  for (k0, b0), (k1, b1) in zip(w, got):
    np.testing.assert_array_equal(k0, k1)
    np.testing.assert_array_equal(b0, b1)


def test_checkpoint_detects_corruption(tmp_path):
  d = str(tmp_path / 'ckpt')
  C.save_conv_weights(d, _weights(2))
  path = os.path.join(d, 'model.ckpt.data-00000-of-00001')
  raw = bytearray(open(path, 'rb').read())
  raw[100] ^= 0x40
  open(path, 'wb').write(bytes(raw))
  with pytest.raises(ValueError, match='CRC'):
    C.load_conv_weights(d)
  C.load_conv_weights(d, verify=False)               # explicit opt-out still reads
  with pytest.raises(ValueError, match='magic'):
    C.read_index(path)


def test_block_prefix_compression():
  # keys sharing prefixes, restart interval 2: the reader must rebuild keys from (shared, unshared)
  items = [(b'predict_coefficients/conv1d/bias', b'A'), (b'predict_coefficients/conv1d/kernel', b'BB'),
           (b'predict_coefficients/conv1d_1/bias', b'CCC')]
  block = C._build_block(items, restart_interval=2)
  assert C._block_entries(block) == items
  assert len(block) < sum(len(k) + len(v) for k, v in items) + 3 * 3 + 12     # compression really happened


def test_num_layers_zero_checkpoint(tmp_path):
  prefix = str(tmp_path / 'model.ckpt')
  coef = np.arange(14, dtype=np.float32).reshape(2, 7)
  C.write_checkpoint(prefix, {'predict_coefficients/coefficients': coef})
  got = C.load_conv_weights(str(tmp_path))
  assert len(got) == 1
  np.testing.assert_array_equal(got[0], coef)


HPARAMS_TEXT = '''
hparam {
  key: "equation"
  value {
    bytes_value: "kdv"
  }
}
hparam {
  key: "conservative"
  value {
    bool_value: false
  }
}
hparam {
  key: "equation_kwargs"
  value {
    bytes_value: "{\\"num_points\\": 128}"
  }
}
hparam {
  key: "learning_rates"
  value {
    float_list {
      value: 0.0010000000474974513
      value: 9.999999747378752e-05
    }
  }
}
hparam {
  key: "learning_stops"
  value {
    int64_list {
      value: 20000
      value: 40000
    }
  }
}
hparam {
  key: "num_layers"
  value {
    int64_value: 4
  }
}
hparam { key: "error_scale" value { float_list { value: nan } } }
hparam { key: "resample_factor" value { int64_value: 8 } }
'''


def test_parse_hparams_pbtxt(tmp_path):
  values = C.parse_hparams_pbtxt(HPARAMS_TEXT)
  assert values['equation'] == 'kdv' and values['conservative'] is False
  assert values['equation_kwargs'] == '{"num_points": 128}'
  assert values['learning_stops'] == [20000, 40000] and values['num_layers'] == 4
  assert values['learning_rates'][0] == pytest.approx(1e-3) and np.isnan(values['error_scale'][0])
  (tmp_path / 'hparams.pbtxt').write_text(HPARAMS_TEXT)
  hp = ddd.training.load_hparams(str(tmp_path))
  assert hp.equation == 'kdv' and hp.num_layers == 4 and hp.resample_factor == 8
  assert hp.kernel_size == 5 and hp.filter_size == 32          # back-filled defaults (training.py:646-647)
  _, coarse = ddd.equations.from_hparams(hp)
  assert coarse.grid.solution_num_points == 128 // 8


def test_hparams_round_trip(tmp_path):
  hp = ddd.training.create_hparams('burgers', conservative=False, num_layers=2, polynomial_accuracy_scale=0.5,
                                   equation_kwargs='{"num_points": 64, "eta": 0.04}', error_scale=[1.5, 2.0])
  ddd.training.save_hparams(str(tmp_path), hp)
  back = ddd.training.load_hparams(str(tmp_path))
  a, b = hp.values(), back.values()
  assert set(a) == set(b)
  for key in a:
    if isinstance(a[key], float) or (isinstance(a[key], list) and a[key] and isinstance(a[key][0], float)):
      np.testing.assert_allclose(np.asarray(b[key], float), np.asarray(a[key], float), rtol=1e-6, equal_nan=True)
    else:
      assert a[key] == b[key], key
