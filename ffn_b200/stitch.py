"""Cross-slab id reconciliation ("stitching") of independently segmented subvolumes.

The reference segments subvolumes independently, each with a private id space, and leaves the assembly of
a global segmentation to the user: "reconciled into a single global ID space ... maintaining a union-find data
structure ... currently *not implemented* in this repository" (doc/manual.md:119-127).  `distributed.merge_slabs`
already makes the ids of all slabs globally unique; this module adds the reconciliation step for slabs that
TOUCH (no overlap, the layout `distributed.slab_boxes` produces): objects on the two sides of a shared face are
the same neurite when their cross-sections on the two face planes coincide.

Criterion (documented, deterministic; there is no reference behaviour to match): for the last plane `A` of one slab
and the first plane `B` of its neighbour, count the voxel pairs (a, b) = (A[p], B[p]) with a > 0 and b > 0.  The pair
is joined when the contact is at least `min_contact` voxels AND at least `min_fraction` of the smaller of the two
cross-sections (|A == a|, |B == b|).  Joined pairs go through a union-find; every set takes its smallest id.

Works on torch tensors on any device (CUDA tensors after the NCCL gather on rank 0, CPU tensors in tests): the
per-face pair counting and the final relabelling are vectorised torch ops, the union-find runs on the host over
the (few) joined pairs.
"""

from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np


class UnionFind:
  """Union-find over arbitrary non-negative integer ids (path halving, union by smaller id)."""

  def __init__(self):
    self.parent: Dict[int, int] = {}

  def find(self, x: int) -> int:
    p = self.parent
    p.setdefault(x, x)
    while p[x] != x:
      p[x] = p[p[x]]
      x = p[x]
    return x

  def union(self, a: int, b: int):
    ra, rb = self.find(a), self.find(b)
    if ra != rb:
      lo, hi = (ra, rb) if ra < rb else (rb, ra)
      self.parent[hi] = lo

  def mapping(self) -> Dict[int, int]:
    """id -> representative (the smallest id of its set), only for ids that changed."""
    return {x: self.find(x) for x in list(self.parent) if self.find(x) != x}


def face_contacts(plane_a, plane_b):
  """(a, b, contact, area_a, area_b) int64 arrays for every label pair touching across two face planes."""
  import torch
  a = plane_a.reshape(-1).to(torch.int64)
  b = plane_b.reshape(-1).to(torch.int64)
  both = (a > 0) & (b > 0)
  if not bool(both.any()):
    z = np.zeros(0, dtype=np.int64)
    return z, z, z, z, z
  base = int(max(int(a.max()), int(b.max()))) + 1
  keys, counts = torch.unique(a[both] * base + b[both], return_counts=True)
  ia, ib = keys // base, keys % base
  ua, ca = torch.unique(a[a > 0], return_counts=True)
  ub, cb = torch.unique(b[b > 0], return_counts=True)
  area_a = ca[torch.searchsorted(ua, ia)]
  area_b = cb[torch.searchsorted(ub, ib)]
  return tuple(t.cpu().numpy() for t in (ia, ib, counts, area_a, area_b))


def stitch_pairs(slabs: Dict[Tuple[int, int, int], "object"], min_contact: int = 16, min_fraction: float = 0.5):
  """Joined (id_a, id_b) pairs over all faces shared by grid-adjacent slabs.

  Args:
    slabs: {(iz, iy, ix) grid index: int label tensor (z, y, x)} with globally unique ids > 0
  """
  pairs = []
  for (iz, iy, ix), lab in slabs.items():
    for axis, nb in ((0, (iz + 1, iy, ix)), (1, (iz, iy + 1, ix)), (2, (iz, iy, ix + 1))):
      other = slabs.get(nb)
      if other is None:
        continue
      plane_a = lab.select(axis, lab.shape[axis] - 1)
      plane_b = other.select(axis, 0)
      if tuple(plane_a.shape) != tuple(plane_b.shape):
        raise ValueError('slabs %r and %r do not share a face of equal size' % ((iz, iy, ix), nb))
      ia, ib, cnt, area_a, area_b = face_contacts(plane_a, plane_b)
      keep = (cnt >= min_contact) & (cnt >= min_fraction * np.minimum(area_a, area_b))
      pairs.extend(zip(ia[keep].tolist(), ib[keep].tolist()))
  return pairs


def relabel_(lab, mapping: Dict[int, int]):
  """Applies {old id: new id} to a label tensor IN PLACE (lookup table on the tensor's device)."""
  import torch
  if not mapping:
    return lab
  top = int(lab.max())
  lut = torch.arange(top + 1, dtype=lab.dtype, device=lab.device)
  olds = [o for o in mapping if o <= top]
  if olds:
    lut[torch.tensor(olds, dtype=torch.int64, device=lab.device)] = torch.tensor(
        [mapping[o] for o in olds], dtype=lab.dtype, device=lab.device)
  pos = lab > 0
  lab[pos] = lut[lab[pos].to(torch.int64)]
  return lab


def stitch_slabs(slabs: Dict[Tuple[int, int, int], "object"], min_contact: int = 16, min_fraction: float = 0.5):
  """Reconciles the ids of touching slabs in place; returns ({old id: new id}, number of joined pairs)."""
  uf = UnionFind()
  pairs = stitch_pairs(slabs, min_contact, min_fraction)
  for a, b in pairs:
    uf.union(int(a), int(b))
  mapping = uf.mapping()
  for lab in slabs.values():
    relabel_(lab, mapping)
  return mapping, len(pairs)


def grid_of(gathered: Sequence[Sequence["object"]], world: int, n_slabs: int):
  """[local slab index][rank] lists of `distributed.merge_slabs` -> {(iz, iy, ix): tensor}, using the slab order of
  `distributed.slab_boxes` (C order of the slab grid) and the contiguous assignment of `slabs_of_rank`."""
  from . import distributed as D
  grid = D.slab_grid(n_slabs)
  out = {}
  for rank in range(world):
    for i, k in enumerate(D.slabs_of_rank(n_slabs, rank, world)):
      iz, rem = divmod(k, grid[1] * grid[2])
      iy, ix = divmod(rem, grid[2])
      out[(iz, iy, ix)] = gathered[i][rank]
  return out
