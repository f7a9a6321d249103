"""Seed policies: iterators over (z, y, x) starting points, computed once per canvas on the host.

Mirrors ffn/inference/seed.py: `BaseSeedPolicy` (:37-130, incl. the border filter :81-88 and the
checkpoint state :100-113), `PolicyPeaks` (:142-199), `PolicyMax` (:307-313), `PolicyGrid3d`
(:411-430), `PolicyGrid2d` (:433-452), `PolicyInvertOrigins` (:455-469), `PolicyDenseSeeds`
(:472-492), `ReverseCoords` (:495-504), `SequentialPolicies` (:507-544).

`PolicyPeaks` depends upstream on `edt.edt` and `skimage.feature.peak_local_max`, neither of
which is installable offline; it is restated with scipy (exact Euclidean distance transform with
anisotropy, maximum-filter peak detection with the same seeded tie-break noise).  Seed ORDER and
the tie-break follow the reference; exact equality with skimage's peak suppression is not
guaranteed, which is why flood-fill parity always feeds both sides the same list.
"""

import weakref

import numpy as np
from scipy import ndimage

from . import storage


class BaseSeedPolicy:
  """Base class: subclasses fill `self.coords` ([N, 3] zyx) in `init_coords`."""

  def __init__(self, canvas, **kwargs):
    del kwargs
    self.canvas = weakref.proxy(canvas)
    self.coords = None
    self.idx = 0

  def init_coords(self):
    raise NotImplementedError()

  def __iter__(self):
    return self

  def _ensure_coords(self):
    """Runs init_coords once and applies the border filter (seed.py:76-88)."""
    if self.coords is None:
      self.init_coords()
      if self.coords is None:
        return False
      self.coords = np.asarray(self.coords).reshape(-1, 3)
      if self.coords.size:
        margin = np.array(self.canvas.margin)[np.newaxis, ...]
        keep = np.all((self.coords - margin >= 0) & (self.coords + margin < self.canvas.shape), axis=1)
        self.coords = self.coords[keep, :]
    return True

  def __next__(self):
    if not self._ensure_coords():
      raise StopIteration()
    if self.idx < self.coords.shape[0]:
      curr = self.coords[self.idx, :]
      self.idx += 1
      return tuple(int(v) for v in curr)
    raise StopIteration()

  def next(self):
    return self.__next__()

  def remaining(self):
    """All not-yet-consumed seeds as an [N, 3] array (what the device loop is given)."""
    if not self._ensure_coords():
      return np.zeros((0, 3), dtype=np.int32)
    return np.ascontiguousarray(self.coords[self.idx:], dtype=np.int32)

  def get_state(self, previous=False):
    if previous:
      return self.coords, max(0, self.idx - 1)
    return self.coords, self.idx

  def set_state(self, state):
    self.coords, self.idx = state

  def get_exclusion_mask(self):
    mask = np.asarray(self.canvas.segmentation) > 0
    if self.canvas.restrictor is not None:
      if self.canvas.restrictor.mask is not None:
        mask |= self.canvas.restrictor.mask
      if self.canvas.restrictor.seed_mask is not None:
        mask |= self.canvas.restrictor.seed_mask
    return mask


def _peak_local_max(values, min_distance, threshold_abs):
  """Local maxima of `values` separated by > min_distance (Chebyshev), best first."""
  size = 2 * min_distance + 1
  footprint_max = ndimage.maximum_filter(values, size=size, mode='nearest')
  peaks = (values == footprint_max) & (values > threshold_abs)
  coords = np.argwhere(peaks)
  if coords.shape[0] == 0:
    return coords
  order = np.argsort(-values[tuple(coords.T)], kind='stable')
  return coords[order]


def _find_peaks(distances, min_distance, threshold_abs=0, threshold_rel=0):
  del threshold_rel
  rng = np.random.RandomState(seed=42)   # seed.py:133-139: reproducible tie-break noise
  return _peak_local_max(distances + rng.rand(*distances.shape) * 1e-4, min_distance, threshold_abs)


_NOISE_CACHE = {}


def _tie_break_noise(shape):
  """RandomState(42).rand(*shape) (seed.py:133-139), cached for the most recent shape: a run over many
  equally sized subvolumes draws the same 8 bytes / voxel every time (0.7 s for 512^3)."""
  if shape not in _NOISE_CACHE:
    _NOISE_CACHE.clear()
    _NOISE_CACHE[shape] = np.random.RandomState(seed=42).rand(*shape)
  return _NOISE_CACHE[shape]


class PolicyPeaks(BaseSeedPolicy):
  """Sobel edges -> adaptive threshold -> distance transform -> local maxima (seed.py:142-199)."""

  def init_coords(self):
    dev = getattr(self.canvas, '_dev', None)
    restrictor = self.canvas.restrictor
    # The device's movement mask also carries the shift-mask rule, which seed.py:170-173 does not apply
    # to the edge map: with a shift mask the seeds are computed by the host restatement below.
    if dev is not None and getattr(restrictor, 'shift_mask', None) is None:
      # device path: the canvas' image / segmentation / masks are already resident in HBM
      noise = _tie_break_noise(tuple(int(v) for v in self.canvas.shape))
      with self.canvas._exec_client.engine_lock:         # pylint: disable=protected-access
        self.coords = dev.seed_peaks(self.canvas.voxel_size_zyx, noise).astype(np.int64).reshape(-1, 3)
      return
    # host path (shift masks / no device canvas): the same stages with scipy
    blocked = self._blocked_voxels()
    is_edge = _edge_mask(np.asarray(self.canvas.image), blocked)
    if is_edge.all():
      return
    dist = _distance_map(is_edge, self.canvas.voxel_size_zyx, self.get_exclusion_mask())
    peaks = _find_peaks(dist, min_distance=3, threshold_abs=0, threshold_rel=0)
    self.coords = np.array(sorted(map(tuple, peaks.astype(int).tolist()))).reshape(-1, 3)

  def _blocked_voxels(self):
    """Voxels the restrictor forbids (movement mask | seed mask), or None."""
    r = self.canvas.restrictor
    parts = [m for m in (getattr(r, 'mask', None), getattr(r, 'seed_mask', None)) if m is not None]
    if not parts:
      return None
    out = np.zeros(self.canvas.shape, dtype=bool)
    for m in parts:
      out |= np.asarray(m).astype(bool)
    return out


def _edge_mask(image, blocked):
  """Sobel gradient magnitude above its local (gaussian, sigma 49/6) average; blocked voxels count as edges so that
  large masked regions do not distort the distance transform (seed.py:152-173)."""
  grad = ndimage.generic_gradient_magnitude(image.astype(np.float32), ndimage.sobel)
  local = np.empty_like(grad)
  ndimage.gaussian_filter(grad, 49.0 / 6.0, output=local, mode='reflect')
  out = grad > local
  if blocked is not None:
    out |= blocked
  return out


def _distance_map(is_edge, voxel_size_zyx, excluded):
  """Distance (physical units) to the nearest edge, -1 where seeds are not allowed (seed.py:183-188)."""
  dist = ndimage.distance_transform_edt(~is_edge, sampling=voxel_size_zyx).astype(np.float32)
  dist[excluded | ~np.isfinite(dist)] = -1
  return dist


class PolicyMax(BaseSeedPolicy):
  """All voxels in descending order of intensity (seed.py:307-313)."""

  def init_coords(self):
    image = np.asarray(self.canvas.image)
    order = np.argsort(image.ravel())[::-1]
    self.coords = np.stack(np.unravel_index(order, image.shape), axis=1)


class PolicyGrid3d(BaseSeedPolicy):
  """Uniform 3d grid visited offset by offset (seed.py:411-430)."""

  def __init__(self, canvas, step=16, offsets=(0, 8, 4, 12, 2, 10, 14), **kwargs):
    super().__init__(canvas, **kwargs)
    self.step = step
    self.offsets = offsets

  def init_coords(self):
    shape = self.canvas.shape
    coords = []
    for offset in self.offsets:
      for z in range(offset, shape[0], self.step):
        for y in range(offset, shape[1], self.step):
          for x in range(offset, shape[2], self.step):
            coords.append((z, y, x))
    self.coords = np.array(coords).reshape(-1, 3)


class PolicyGrid2d(BaseSeedPolicy):
  """Uniform 2d grid on every z plane (seed.py:433-452)."""

  def __init__(self, canvas, step=16, offsets=(0, 8, 4, 12, 2, 6, 10, 14), **kwargs):
    super().__init__(canvas, **kwargs)
    self.step = step
    self.offsets = offsets

  def init_coords(self):
    shape = self.canvas.shape
    coords = []
    for offset in self.offsets:
      for z in range(shape[0]):
        for y in range(offset, shape[1], self.step):
          for x in range(offset, shape[2], self.step):
            coords.append((z, y, x))
    self.coords = np.array(coords).reshape(-1, 3)


class PolicyInvertOrigins(BaseSeedPolicy):
  """Seeds of a previous run in reverse id order (seed.py:455-469)."""

  def __init__(self, canvas, corner=None, segmentation_dir=None, **kwargs):
    super().__init__(canvas, **kwargs)
    self.corner = corner
    self.segmentation_dir = segmentation_dir

  def init_coords(self):
    origins = storage.load_origins(self.segmentation_dir, self.corner)
    points = sorted(origins.items(), reverse=True)
    self.coords = np.array([info.start_zyx for _, info in points]).reshape(-1, 3)


class PolicyDenseSeeds(BaseSeedPolicy):
  """Every voxel above a threshold, after optional erosions (seed.py:472-492)."""

  def __init__(self, canvas, threshold=0.5, num_erosions=0, invert=False, **kwargs):
    super().__init__(canvas, **kwargs)
    self._threshold = threshold
    self._num_erosions = num_erosions
    self._invert = invert

  def init_coords(self):
    x = np.asarray(self.canvas.image) > self._threshold
    if self._invert:
      x = ~x
    for _ in range(self._num_erosions):
      # skimage.morphology.binary_erosion (seed.py:488): connectivity-1 cross, outside of the image counts as foreground
      x = ndimage.binary_erosion(x, border_value=True)
    self.coords = np.array(np.where(x)).T


class ReverseCoords(BaseSeedPolicy):
  """Wraps another policy and reverses its order (seed.py:495-504)."""

  def __init__(self, canvas, policy_to_reverse, **policy_kwargs):
    super().__init__(canvas)
    self._policy = globals()[policy_to_reverse](canvas, **policy_kwargs)

  def init_coords(self):
    self.coords = np.array(list(self._policy)[::-1]).reshape(-1, 3)


class SequentialPolicies(BaseSeedPolicy):
  """Runs several policies one after another (seed.py:507-544)."""

  def __init__(self, canvas, policies, **kwargs):
    super().__init__(canvas, **kwargs)
    self._policies = [globals()[name](canvas, **dict(args, **kwargs)) for name, args in policies]

  def init_coords(self):
    chunks = [np.array(list(p)).reshape(-1, 3) for p in self._policies]
    self.coords = np.concatenate(chunks, axis=0) if chunks else np.zeros((0, 3), dtype=np.int64)
