"""A machine-independent stand-in network for pinning the flood-fill *logic* (test infra).

The real conv stack's float32 results depend on the CPU's conv3d kernel selection, so a
trajectory recorded on one machine can differ at knife-edge decisions on another.  This toy
"network" uses only elementwise IEEE float32 operations (+, -, *, abs, clip), which numpy
evaluates identically everywhere.  It grows an object over voxels whose intensity is close to
the intensity at the FoV centre — enough to exercise moves in all six directions, the BFS queue,
the quantised done-set, the disco merge, overlaps and the small/weak-seed rejections.
"""

import numpy as np


def toy_net(seed: np.ndarray, image: np.ndarray) -> np.ndarray:
  seed = np.asarray(seed, dtype=np.float32)
  image = np.asarray(image, dtype=np.float32)
  c = image[tuple(s // 2 for s in image.shape)]
  sim = np.float32(1.0) - np.abs(image - c) * np.float32(4.0)
  target = np.clip(sim * np.float32(6.0) - np.float32(2.5), np.float32(-4.0), np.float32(4.0))
  return (seed + (target - seed) * np.float32(0.875)).astype(np.float32)


def toy_image(cells: np.ndarray) -> np.ndarray:
  """Piecewise-constant float32 image from integer cell ids (0 = membrane)."""
  z, y, x = np.indices(cells.shape)
  level = ((cells.astype(np.int64) * 37) % 11).astype(np.float32) / np.float32(5.0) - np.float32(1.0)
  ramp = (((x * 7 + y * 13 + z * 29) % 17).astype(np.float32) / np.float32(17.0)) * np.float32(0.004)
  img = np.where(cells > 0, level + ramp, np.float32(-2.0)).astype(np.float32)
  return img
