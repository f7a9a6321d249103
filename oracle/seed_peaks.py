"""Oracle: the PolicyPeaks seed list on the CPU (TEST INFRASTRUCTURE — only tests/, smoke() and bench.py's
CPU-baseline leg may import this; the product never does).

Restates ffn/inference/seed.py:133-199 (`_find_peaks`, `PolicyPeaks.init_coords`), :107-118
(`get_exclusion_mask`) and the border filter of `BaseSeedPolicy.__next__` (:74-88), independently of
`ffn_b200` (nothing from the product package is imported here).

PINNED against the reference's own, unmodified `PolicyPeaks` (tests/golden/make_golden_peaks.py ->
policy_peaks_ref.npz, tests/test_oracle_golden.py::test_seed_peaks_oracle_equals_reference_policy_peaks: isotropic,
anisotropic, masked canvases; the seed lists are equal, order included) — for everything except the two third-party
calls below, which that fixture injects by definition.

Third-party pieces of the reference that are NOT under /root/reference (un-vendored, unpinned in
setup.py:43,47 — `edt>=2.3.0`, `scikit-image>=0.11.0`), so for these two definitions the oracle is **parity unpinned**:

  * `edt.edt(binary, anisotropy=voxel_size)`: the exact Euclidean distance transform of the foreground to the
    nearest background voxel, distances in physical units, float32 output (edt's documented semantics).
    Restated with `scipy.ndimage.distance_transform_edt(sampling=...)` (exact as well, float64) cast to
    float32 — `brute_force_edt` below is the O(n^2) definition, used by the tests to pin the restatement on
    small volumes.  One documented difference is NOT reproduced: edt.edt treats the volume border as
    background only with `black_border=True` (default False), like scipy.
  * `skimage.feature.peak_local_max(image, min_distance=3, threshold_abs=0, threshold_rel=0)`: documented
    algorithm — a voxel is a peak iff it equals the maximum of its (2*min_distance+1)^3 neighbourhood and
    exceeds the threshold; peaks closer than min_distance to the border are dropped (`exclude_border=True`);
    returned best first.  The reference adds RandomState(42).rand * 1e-4 so that the maximum is unique, and
    re-sorts the result lexicographically (seed.py:193-196), so the skimage return order does not matter.
"""

from __future__ import annotations

import numpy as np
from scipy import ndimage


def brute_force_edt(foreground: np.ndarray, voxel_size_zyx=(1.0, 1.0, 1.0)) -> np.ndarray:
  """Definition of the exact EDT, O(n * background) — small volumes only."""
  fg = np.asarray(foreground).astype(bool)
  out = np.zeros(fg.shape, dtype=np.float64)
  bg = np.argwhere(~fg).astype(np.float64) * np.asarray(voxel_size_zyx, dtype=np.float64)[None]
  if bg.shape[0] == 0:
    out[...] = np.inf
    return out
  pts = np.argwhere(fg)
  scaled = pts.astype(np.float64) * np.asarray(voxel_size_zyx, dtype=np.float64)[None]
  for i in range(pts.shape[0]):
    d2 = ((bg - scaled[i][None]) ** 2).sum(axis=1)
    out[tuple(pts[i])] = np.sqrt(d2.min())
  return out


def edge_map(image_f32: np.ndarray) -> np.ndarray:
  """seed.py:152-160: Sobel gradient magnitude above its own gaussian (sigma 49/6) average."""
  edges = ndimage.generic_gradient_magnitude(np.asarray(image_f32).astype(np.float32), ndimage.sobel)
  thresh = np.zeros(edges.shape, dtype=np.float32)
  ndimage.gaussian_filter(edges, 49.0 / 6.0, output=thresh, mode='reflect')
  return edges > thresh


def peak_local_max(values: np.ndarray, min_distance: int, threshold_abs: float) -> np.ndarray:
  """Documented skimage semantics for a field with a unique maximum per neighbourhood; rows (z, y, x)."""
  size = 2 * min_distance + 1
  is_max = values == ndimage.maximum_filter(values, size=size, mode='nearest')
  is_max &= values > threshold_abs
  border = np.zeros(values.shape, dtype=bool)                       # exclude_border=True -> min_distance
  border[tuple(slice(min_distance, s - min_distance) for s in values.shape)] = True
  return np.argwhere(is_max & border)


def policy_peaks(image_f32: np.ndarray, voxel_size_zyx=(1, 1, 1), segmentation=None, mask=None, seed_mask=None,
                 margin_zyx=None) -> np.ndarray:
  """The seed list PolicyPeaks yields, in order: [N, 3] int64 (z, y, x)."""
  filt_edges = edge_map(image_f32)
  excl = np.zeros(filt_edges.shape, dtype=bool) if segmentation is None else (np.asarray(segmentation) > 0)
  if mask is not None:                                               # seed.py:107-118, :170-173
    excl |= np.asarray(mask).astype(bool)
    filt_edges[np.asarray(mask).astype(bool)] = True
  if seed_mask is not None:
    excl |= np.asarray(seed_mask).astype(bool)
    filt_edges[np.asarray(seed_mask).astype(bool)] = True
  if np.all(filt_edges):                                             # seed.py:176-177
    return np.zeros((0, 3), dtype=np.int64)
  dt = ndimage.distance_transform_edt(~filt_edges, sampling=tuple(float(v) for v in voxel_size_zyx)).astype(np.float32)
  dt[excl] = -1
  dt[~np.isfinite(dt)] = -1
  noise = np.random.RandomState(seed=42).rand(*dt.shape)            # seed.py:133-139
  idxs = peak_local_max(dt + noise * 1e-4, min_distance=3, threshold_abs=0)
  coords = np.array(sorted((int(z), int(y), int(x)) for z, y, x in idxs), dtype=np.int64).reshape(-1, 3)
  if margin_zyx is not None and coords.size:                         # seed.py:81-88
    m = np.asarray(margin_zyx)[None]
    coords = coords[np.all((coords - m >= 0) & (coords + m < np.asarray(dt.shape)[None]), axis=1)]
  return coords
