"""Identity alignment (ffn/inference/align.py:20-172): the only alignment the reference ships."""

import numpy as np


class Alignment:
  """Null alignment of a subvolume: every transform is the identity, crops are plain slicing."""

  def __init__(self, corner, size):
    self._corner = np.asarray(corner)
    self._size = np.asarray(size)

  @property
  def corner(self):
    return self._corner

  @property
  def size(self):
    return self._size

  def expand_bounds(self, corner, size, forward=True):
    del forward
    return np.asarray(corner), np.asarray(size)

  def transform_shift_mask(self, corner, scale, mask):
    del corner, scale
    return mask

  def align_and_crop(self, src_corner, source, dst_corner, dst_size, fill=0, forward=True):
    """Crops (and zero/`fill`-pads) `source` located at `src_corner` to the destination box."""
    del forward
    src_corner = np.asarray(src_corner)
    dst_corner = np.asarray(dst_corner)
    dst_size = np.asarray(dst_size)
    src_size = np.asarray(source.shape)
    if np.all(src_corner == dst_corner) and np.all(src_size == dst_size):
      return source
    out = np.full(tuple(int(v) for v in dst_size), fill, dtype=source.dtype)
    lo = np.maximum(src_corner, dst_corner)
    hi = np.minimum(src_corner + src_size, dst_corner + dst_size)
    if np.all(hi > lo):
      s = tuple(slice(int(a), int(b)) for a, b in zip(lo - src_corner, hi - src_corner))
      d = tuple(slice(int(a), int(b)) for a, b in zip(lo - dst_corner, hi - dst_corner))
      out[d] = source[s]
    return out

  def transform(self, zyx, forward=True):
    del forward
    return zyx

  def rescaled(self, zyx_scale):
    zyx_scale = np.asarray(zyx_scale)
    return Alignment(zyx_scale * self.corner, zyx_scale * self.size)


class Aligner:
  """Factory of identity alignments (align.py:153-172)."""

  def generate_alignment(self, corner, size):
    return Alignment(corner, size)
