"""Axis-aligned integer boxes in (x, y, z) order, interoperable with the BoundingBox proto.

The subset of ffn/utils/bounding_box.py:29-185 that the inference path touches (`start` / `size` /
`end`, `to_slice`, `to_proto`, `adjusted_by`, `intersection`, `containing`), written for this package.
"""

import numpy as np

from . import bounding_box_pb2


def _vec3(v):
  """Sequence, numpy array or Vector3j-like proto -> int64 numpy (x, y, z)."""
  if hasattr(v, 'x') and hasattr(v, 'y') and hasattr(v, 'z'):
    return np.array([v.x, v.y, v.z], dtype=np.int64)
  arr = np.asarray(v, dtype=np.int64).reshape(-1)
  if arr.shape != (3,):
    raise ValueError('expected 3 components, got %r' % (v,))
  return arr


class BoundingBox:
  """Box given by exactly two of start / size / end (end exclusive), or copied from a box / proto."""

  def __init__(self, start=None, size=None, end=None):
    if start is not None and hasattr(start, 'start') and hasattr(start, 'size'):
      if size is not None or end is not None:
        raise ValueError('a BoundingBox object/proto must be specified alone')
      start, size = start.start, start.size
    if sum(v is not None for v in (start, size, end)) != 2:
      raise ValueError('exactly two of start, end, and size must be specified')
    if start is None:
      self.size = _vec3(size)
      self.start = _vec3(end) - self.size
    elif size is None:
      self.start = _vec3(start)
      self.size = _vec3(end) - self.start
    else:
      self.start, self.size = _vec3(start), _vec3(size)

  @property
  def end(self):
    return self.start + self.size

  def adjusted_by(self, start=None, end=None):
    """New box with the bounds moved by the given (x, y, z) amounts."""
    lo = self.start + (_vec3(start) if start is not None else 0)
    hi = self.end + (_vec3(end) if end is not None else 0)
    return BoundingBox(start=lo, end=hi)

  def to_proto(self):
    proto = bounding_box_pb2.BoundingBox()
    for name, vec in (('start', self.start), ('size', self.size)):
      field = getattr(proto, name)
      field.x, field.y, field.z = (int(v) for v in vec)
    return proto

  def to_slice(self):
    """Index expression in array (z, y, x) order."""
    lo, hi = self.start, self.end
    return np.index_exp[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]]

  def __repr__(self):
    return 'BoundingBox(start=%s, size=%s)' % (tuple(int(v) for v in self.start), tuple(int(v) for v in self.size))

  def __eq__(self, other):
    if not isinstance(other, BoundingBox):
      other = BoundingBox(other)
    return bool(np.all(self.start == other.start) and np.all(self.size == other.size))

  def __ne__(self, other):
    return not self == other

  def __hash__(self):
    return hash((tuple(int(v) for v in self.start), tuple(int(v) for v in self.size)))


def intersection(box0, box1):
  """The overlap of two boxes, or None when they do not intersect."""
  lo = np.maximum(box0.start, box1.start)
  hi = np.minimum(box0.end, box1.end)
  if np.any(hi <= lo):
    return None
  return BoundingBox(start=lo, end=hi)


def containing(*boxes):
  """The smallest box containing all the given ones."""
  if not boxes:
    raise ValueError('at least one box is required')
  lo = np.min([b.start for b in boxes], axis=0)
  hi = np.max([b.end for b in boxes], axis=0)
  return BoundingBox(start=lo, end=hi)
