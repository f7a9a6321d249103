"""TensorFlow-free reader for TF1 "bundle" checkpoints (``model.ckpt-N.{index,data-*}``).

The reference restores its weights with ``tf.train.Saver().restore``
(ffn/inference/runner.py:98-111).  TensorFlow is not part of this framework, so the two files are
parsed directly:

* ``.index`` is a leveldb-format table (footer -> index block -> data blocks with prefix-compressed
  keys); every key is a variable name and every value a serialized ``BundleEntryProto``
  {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}.
* ``.data-00000-of-00001`` holds the raw little-endian tensor bytes at ``offset``.

Only what the FFN inference path needs is implemented: uncompressed blocks, DT_FLOAT / DT_INT32 /
DT_INT64 tensors, a single shard.
"""

from __future__ import annotations

import os
import struct
from typing import Dict, Tuple

import numpy as np

_TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
  result = 0
  shift = 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def _read_block(data: bytes, offset: int, size: int) -> bytes:
  block = data[offset:offset + size]
  ctype = data[offset + size]
  if ctype != 0:
    raise ValueError('compressed checkpoint index blocks are not supported (type %d)' % ctype)
  return block


def _block_entries(block: bytes):
  """Yields (key, value) pairs of one table block (prefix-compressed keys)."""
  num_restarts = struct.unpack('<I', block[-4:])[0]
  limit = len(block) - 4 - 4 * num_restarts
  pos = 0
  key = b''
  while pos < limit:
    shared, pos = _varint(block, pos)
    non_shared, pos = _varint(block, pos)
    value_len, pos = _varint(block, pos)
    key = key[:shared] + block[pos:pos + non_shared]
    pos += non_shared
    value = block[pos:pos + value_len]
    pos += value_len
    yield key, value


def _parse_shape(buf: bytes):
  dims = []
  pos = 0
  while pos < len(buf):
    tag, pos = _varint(buf, pos)
    field, wire = tag >> 3, tag & 7
    if wire == 2:
      ln, pos = _varint(buf, pos)
      sub = buf[pos:pos + ln]
      pos += ln
      if field == 2:  # TensorShapeProto.Dim
        size = 0
        sp = 0
        while sp < len(sub):
          stag, sp = _varint(sub, sp)
          if stag & 7 == 0:
            val, sp = _varint(sub, sp)
            if stag >> 3 == 1:
              size = val
          elif stag & 7 == 2:
            sl, sp = _varint(sub, sp)
            sp += sl
          else:
            raise ValueError('unexpected wire type in TensorShapeProto.Dim')
        dims.append(size)
    elif wire == 0:
      _, pos = _varint(buf, pos)
    else:
      raise ValueError('unexpected wire type in TensorShapeProto')
  return tuple(dims)


def _parse_entry(buf: bytes) -> dict:
  out = {'dtype': 0, 'shape': (), 'shard_id': 0, 'offset': 0, 'size': 0}
  pos = 0
  while pos < len(buf):
    tag, pos = _varint(buf, pos)
    field, wire = tag >> 3, tag & 7
    if wire == 0:
      val, pos = _varint(buf, pos)
      if field == 1:
        out['dtype'] = val
      elif field == 3:
        out['shard_id'] = val
      elif field == 4:
        out['offset'] = val
      elif field == 5:
        out['size'] = val
    elif wire == 2:
      ln, pos = _varint(buf, pos)
      if field == 2:
        out['shape'] = _parse_shape(buf[pos:pos + ln])
      pos += ln
    elif wire == 5:
      pos += 4
    elif wire == 1:
      pos += 8
    else:
      raise ValueError('unexpected wire type %d in BundleEntryProto' % wire)
  return out


def list_variables(prefix: str) -> Dict[str, dict]:
  """Returns {variable name: {dtype, shape, shard_id, offset, size}} for checkpoint `prefix`."""
  with open(prefix + '.index', 'rb') as f:
    data = f.read()
  if len(data) < 48:
    raise ValueError('not a TF checkpoint index: %s.index' % prefix)
  footer = data[-48:]
  if struct.unpack('<Q', footer[-8:])[0] != _TABLE_MAGIC:
    raise ValueError('bad table magic in %s.index' % prefix)
  pos = 0
  _, pos = _varint(footer, pos)  # metaindex offset
  _, pos = _varint(footer, pos)  # metaindex size
  index_off, pos = _varint(footer, pos)
  index_size, pos = _varint(footer, pos)

  entries = {}
  for _, handle in _block_entries(_read_block(data, index_off, index_size)):
    boff, hp = _varint(handle, 0)
    bsize, hp = _varint(handle, hp)
    for key, value in _block_entries(_read_block(data, boff, bsize)):
      if not key:  # header entry (BundleHeaderProto)
        continue
      entries[key.decode('utf-8')] = _parse_entry(value)
  return entries


def load_variables(prefix: str) -> Dict[str, np.ndarray]:
  """Reads every tensor of checkpoint `prefix` into a dict of numpy arrays."""
  entries = list_variables(prefix)
  shard_paths = sorted(
      p for p in os.listdir(os.path.dirname(prefix) or '.')
      if p.startswith(os.path.basename(prefix) + '.data-'))
  if len(shard_paths) != 1:
    raise ValueError('expected exactly one data shard for %s, found %r' % (prefix, shard_paths))
  with open(os.path.join(os.path.dirname(prefix) or '.', shard_paths[0]), 'rb') as f:
    blob = f.read()
  out = {}
  for name, e in entries.items():
    if e['dtype'] not in _DTYPES:
      continue
    dt = np.dtype(_DTYPES[e['dtype']]).newbyteorder('<')
    n = int(np.prod(e['shape'])) if e['shape'] else 1
    arr = np.frombuffer(blob, dtype=dt, count=n, offset=e['offset'])
    if n * dt.itemsize != e['size']:
      raise ValueError('size mismatch for %s' % name)
    out[name] = arr.reshape(e['shape']).astype(dt.newbyteorder('='), copy=True)
  return out


def load_convstack_weights(prefix: str, depth: int, scope: str = 'seed_update'):
  """Returns (weights, biases) lists for ConvStack3DFFNModel in layer order.

  Layer order follows ffn/training/models/convstack_3d.py:38-54: conv0_a, conv0_b,
  conv1_a, conv1_b, ..., conv{depth-1}_b, conv_lom.  Weights stay in TF's DHWIO layout
  ``[kz, ky, kx, cin, cout]``.
  """
  v = load_variables(prefix)
  names = []
  for i in range(depth):
    names += ['conv%d_a' % i, 'conv%d_b' % i]
  names.append('conv_lom')
  weights, biases = [], []
  for n in names:
    weights.append(np.ascontiguousarray(v['%s/%s/weights' % (scope, n)], dtype=np.float32))
    biases.append(np.ascontiguousarray(v['%s/%s/biases' % (scope, n)], dtype=np.float32))
  return weights, biases


def load_convstack_npz(path: str):
  """Reads (weights, biases) saved as w00..wNN / b00..bNN arrays (DHWIO), layer order as above."""
  z = np.load(path)
  n = len([k for k in z.files if k.startswith('w')])
  return ([np.ascontiguousarray(z['w%02d' % i], dtype=np.float32) for i in range(n)],
          [np.ascontiguousarray(z['b%02d' % i], dtype=np.float32) for i in range(n)])
