"""Multi-GPU orchestration: one process per GPU, disjoint subvolumes ("slabs"), one merge at the end.

The reference has no multi-process inference (doc/manual.md:107-127): subvolumes are independent
and every subvolume has a private id space.  Here each rank flood-fills its own slabs with its
own engine — no collective on the data path — and the only exchange step is the final merge:

  1. all_gather of each rank's max id (one int64)  -> exclusive prefix sum = id offset per rank
  2. `ffn_canvas_add_id_offset` (HBM-bound kernel) makes ids globally unique
  3. gather / all_gather of the int32 label slabs (and uint8 probability slabs) over NCCL

With the `gloo` backend and CPU tensors the same code runs without GPUs (tests).
"""

from __future__ import annotations

import numpy as np


def slab_grid(n_slabs: int):
  """Factorises n_slabs (power of two) into a (z, y, x) grid, splitting z first."""
  grid = [1, 1, 1]
  axis = 0
  n = n_slabs
  while n > 1:
    if n % 2:
      raise ValueError('number of slabs must be a power of two')
    grid[axis] *= 2
    n //= 2
    axis = (axis + 1) % 3
  return tuple(grid)


def slab_boxes(shape_zyx, n_slabs: int):
  """Disjoint (corner, size) boxes covering the volume, in C order of the slab grid."""
  grid = slab_grid(n_slabs)
  edges = [np.linspace(0, s, g + 1).astype(int) for s, g in zip(shape_zyx, grid)]
  boxes = []
  for iz in range(grid[0]):
    for iy in range(grid[1]):
      for ix in range(grid[2]):
        lo = (edges[0][iz], edges[1][iy], edges[2][ix])
        hi = (edges[0][iz + 1], edges[1][iy + 1], edges[2][ix + 1])
        boxes.append((tuple(int(v) for v in lo), tuple(int(h - l) for h, l in zip(hi, lo))))
  return boxes


def slabs_of_rank(n_slabs: int, rank: int, world: int):
  """Round-robin-free contiguous assignment: rank r owns slabs [r*k, (r+1)*k)."""
  per = -(-n_slabs // world)
  return list(range(rank * per, min((rank + 1) * per, n_slabs)))


def exclusive_offsets(max_ids):
  """[m0, m1, ...] -> [0, m0, m0+m1, ...]."""
  out, acc = [], 0
  for m in max_ids:
    out.append(acc)
    acc += int(m)
  return out


def gather_max_ids(local_max_id: int, device=None):
  """all_gather of one int64 per rank (NCCL on `device`, gloo on CPU)."""
  import torch
  import torch.distributed as dist
  world = dist.get_world_size()
  t = torch.tensor([int(local_max_id)], dtype=torch.int64, device=device)
  outs = [torch.zeros_like(t) for _ in range(world)]
  dist.all_gather(outs, t)
  return [int(o.item()) for o in outs]


class _CudaView:
  """Minimal __cuda_array_interface__ wrapper so torch can alias engine-owned device memory."""

  def __init__(self, ptr, shape, typestr):
    self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False),
                                     'version': 3, 'strides': None}


def canvas_tensor(device_canvas, which, shape, typestr, device):
  """torch tensor aliasing a canvas array in HBM (no copy)."""
  import torch
  ptr, _ = device_canvas.device_ptr(which)
  return torch.as_tensor(_CudaView(ptr, shape, typestr), device=device)


def merge_labels(local_labels, local_max_id: int, dst: int = 0, add_offset=None):
  """Makes ids globally unique and gathers every rank's label slab on `dst`.

  Args:
    local_labels: torch tensor (CPU for gloo, CUDA for NCCL) of this rank's labels, any int dtype
    local_max_id: largest id used by this rank
    dst: rank that receives the list of slabs
    add_offset: optional callable(offset) applying the offset in place on the device (the engine's
      relabel kernel); default adds it with a torch expression.

  Returns:
    (list of tensors on dst | None, this rank's offset, total number of ids)
  """
  import torch
  import torch.distributed as dist
  rank, world = dist.get_rank(), dist.get_world_size()
  max_ids = gather_max_ids(local_max_id, device=local_labels.device if local_labels.is_cuda else None)
  offsets = exclusive_offsets(max_ids)
  off = offsets[rank]
  if off:
    if add_offset is not None:
      add_offset(off)
    else:
      local_labels[local_labels > 0] += off
  gathered = None
  if rank == dst:
    gathered = [torch.empty_like(local_labels) for _ in range(world)]
  dist.gather(local_labels, gathered, dst=dst)
  return gathered, off, sum(max_ids)


def merge_slabs(labels, probs, max_ids, dst: int = 0, add_offset=None):
  """The exchange step of a sharded run: every rank holds the results of its slabs
  (`labels[i]` int32, `probs[i]` uint8 or None, `max_ids[i]` = ids used by slab i); ids are made globally unique
  in slab order (rank-major) and both arrays of every slab are gathered on `dst`.

  Args:
    labels / probs: lists of torch tensors (CUDA for NCCL, CPU for gloo); every rank holds the same number
    max_ids: list of ints, one per local slab
    add_offset: optional callable(i, offset) applying the offset to local slab i in place (the engine's relabel
      kernel); default: a torch expression on labels[i]

  Returns:
    (gathered_labels, gathered_probs, offsets of the local slabs, total ids); the gathered lists are
    [local slab index][rank] on dst and None elsewhere.
  """
  import torch
  import torch.distributed as dist
  rank, world = dist.get_rank(), dist.get_world_size()
  dev = labels[0].device if labels and labels[0].is_cuda else None
  per_rank = gather_max_ids(sum(int(m) for m in max_ids), device=dev)
  off = exclusive_offsets(per_rank)[rank]
  offsets = []
  for i, m in enumerate(max_ids):
    offsets.append(off)
    if off:
      if add_offset is not None:
        add_offset(i, off)
      else:
        labels[i][labels[i] > 0] += off
    off += int(m)
  out_l, out_p = ([] if rank == dst else None), ([] if rank == dst else None)
  for i in range(len(labels)):
    gl = [torch.empty_like(labels[i]) for _ in range(world)] if rank == dst else None
    dist.gather(labels[i], gl, dst=dst)
    if rank == dst:
      out_l.append(gl)
    if probs is not None:
      gp = [torch.empty_like(probs[i]) for _ in range(world)] if rank == dst else None
      dist.gather(probs[i], gp, dst=dst)
      if rank == dst:
        out_p.append(gp)
  return out_l, (out_p if probs is not None else None), offsets, sum(per_rank)
