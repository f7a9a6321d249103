"""Label-volume helpers used when saving results (subset of ffn/inference/segmentation.py)."""

import numpy as np


def reduce_id_bits(segmentation: np.ndarray) -> np.ndarray:
  """Smallest unsigned dtype holding every id (ffn/inference/segmentation.py:66-86)."""
  max_id = segmentation.max() if segmentation.size else 0
  for dt in (np.uint8, np.uint16, np.uint32):
    if max_id <= np.iinfo(dt).max:
      return segmentation.astype(dt)
  return segmentation


def clear_dust(data: np.ndarray, min_size: int = 10) -> np.ndarray:
  """Zeroes segments smaller than `min_size` voxels, in place (segmentation.py:21-63)."""
  ids, sizes = np.unique(data, return_counts=True)
  small = ids[sizes < min_size]
  if small.size:
    data[np.isin(data, small)] = 0
  return data
