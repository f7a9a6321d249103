"""world_size-2 gloo test of the multi-GPU merge logic (runs on CPU)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ffn_b200 import distributed


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out_dir):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  rng = np.random.RandomState(rank)
  local_max = 3 + 2 * rank
  labels = torch.from_numpy(rng.randint(-1, local_max + 1, size=(4, 5, 6)).astype(np.int32))
  before = labels.clone()
  gathered, off, total = distributed.merge_labels(labels, local_max, dst=0)
  np.save(os.path.join(out_dir, 'before_%d.npy' % rank), before.numpy())
  if rank == 0:
    np.save(os.path.join(out_dir, 'merged.npy'), np.stack([g.numpy() for g in gathered]))
    np.save(os.path.join(out_dir, 'total.npy'), np.array([total]))
  dist.destroy_process_group()


def test_merge_labels_two_ranks(tmp_path):
  port = _free_port()
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  merged = np.load(tmp_path / 'merged.npy')
  b0, b1 = np.load(tmp_path / 'before_0.npy'), np.load(tmp_path / 'before_1.npy')
  assert int(np.load(tmp_path / 'total.npy')[0]) == 3 + 5
  np.testing.assert_array_equal(merged[0], b0)                       # rank 0: offset 0
  want1 = b1.copy()
  want1[b1 > 0] += 3                                                 # rank 1: offset = max id of rank 0
  np.testing.assert_array_equal(merged[1], want1)
  ids0 = set(np.unique(merged[0][merged[0] > 0]))
  ids1 = set(np.unique(merged[1][merged[1] > 0]))
  assert not ids0 & ids1


def _worker_slabs(rank, world, port, out_dir):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  rng = np.random.RandomState(10 + rank)
  max_ids = [2 + rank, 4]                       # two slabs per rank
  labels = [torch.from_numpy(rng.randint(-1, m + 1, size=(3, 4, 5)).astype(np.int32)) for m in max_ids]
  probs = [torch.from_numpy(rng.randint(0, 256, size=(3, 4, 5)).astype(np.uint8)) for _ in max_ids]
  np.save(os.path.join(out_dir, 'lab_%d.npy' % rank), np.stack([t.numpy().copy() for t in labels]))
  np.save(os.path.join(out_dir, 'prob_%d.npy' % rank), np.stack([t.numpy().copy() for t in probs]))
  gl, gp, offsets, total = distributed.merge_slabs(labels, probs, max_ids, dst=0)
  np.save(os.path.join(out_dir, 'off_%d.npy' % rank), np.array(offsets))
  if rank == 0:
    np.save(os.path.join(out_dir, 'gl.npy'), np.stack([np.stack([t.numpy() for t in row]) for row in gl]))
    np.save(os.path.join(out_dir, 'gp.npy'), np.stack([np.stack([t.numpy() for t in row]) for row in gp]))
    np.save(os.path.join(out_dir, 'total.npy'), np.array([total]))
  dist.destroy_process_group()


def test_merge_slabs_labels_and_probabilities_two_ranks(tmp_path):
  """configs[3]'s exchange step on CPU: slabs of two ranks, ids unique in rank-major slab order, labels AND
  probability maps gathered on rank 0 unchanged."""
  port = _free_port()
  mp.spawn(_worker_slabs, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  gl, gp = np.load(tmp_path / 'gl.npy'), np.load(tmp_path / 'gp.npy')        # [slab][rank]
  assert int(np.load(tmp_path / 'total.npy')[0]) == (2 + 4) + (3 + 4)
  assert np.load(tmp_path / 'off_0.npy').tolist() == [0, 2] and np.load(tmp_path / 'off_1.npy').tolist() == [6, 9]
  seen = set()
  for r in range(2):
    lab, prob = np.load(tmp_path / ('lab_%d.npy' % r)), np.load(tmp_path / ('prob_%d.npy' % r))
    offs = np.load(tmp_path / ('off_%d.npy' % r))
    for i in range(2):
      want = lab[i].copy()
      want[lab[i] > 0] += offs[i]
      np.testing.assert_array_equal(gl[i, r], want)
      np.testing.assert_array_equal(gp[i, r], prob[i])
      ids = set(np.unique(want[want > 0]).tolist())
      assert not ids & seen
      seen |= ids


def _label_pieces(sub):
  """Connected pieces of every cell of a ground-truth crop, ids from 1 (a slab's private id space)."""
  from scipy import ndimage
  lab = np.zeros(sub.shape, dtype=np.int32)
  nxt = 0
  for cid in np.unique(sub[sub > 0]):
    comp, n = ndimage.label(sub == cid)
    lab[comp > 0] = comp[comp > 0] + nxt
    nxt += n
  return lab, nxt


def _worker_stitch(rank, world, port, out_dir):
  from ffn_b200 import stitch
  from ffn_b200.synthetic import voronoi_phantom
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  _, cells = voronoi_phantom((48, 56, 64), seed=11, return_cells=True, cell_volume=6000.0)
  boxes = distributed.slab_boxes(cells.shape, 8)
  labels, max_ids = [], []
  for k in distributed.slabs_of_rank(8, rank, world):
    lo, size = boxes[k]
    lab, n = _label_pieces(cells[lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]])
    labels.append(torch.from_numpy(lab))
    max_ids.append(n)
  gl, _, _, total = distributed.merge_slabs(labels, None, max_ids, dst=0)
  if rank == 0:
    slabs = stitch.grid_of(gl, world, 8)
    mapping, n_pairs = stitch.stitch_slabs(slabs, min_contact=4, min_fraction=0.5)
    out = np.zeros(cells.shape, dtype=np.int64)
    grid = distributed.slab_grid(8)
    for k, (lo, size) in enumerate(boxes):
      iz, rem = divmod(k, grid[1] * grid[2])
      iy, ix = divmod(rem, grid[2])
      out[lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]] = slabs[(iz, iy, ix)].numpy()
    np.save(os.path.join(out_dir, 'stitched.npy'), out)
    np.save(os.path.join(out_dir, 'cells.npy'), cells)
    np.save(os.path.join(out_dir, 'meta.npy'), np.array([total, n_pairs, len(mapping)]))
  dist.destroy_process_group()


def test_merge_then_stitch_two_ranks(tmp_path):
  """configs[3] exchange step + cross-slab reconciliation (ffn_b200/stitch.py): 8 touching slabs on 2 ranks, private id
  spaces -> NCCL-style gather (gloo here) -> union-find over the shared faces on rank 0.  No stitched object may span
  two true cells, and the cut cells come back as one object."""
  port = _free_port()
  mp.spawn(_worker_stitch, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  out, cells = np.load(tmp_path / 'stitched.npy'), np.load(tmp_path / 'cells.npy')
  total, n_pairs, n_mapped = np.load(tmp_path / 'meta.npy').tolist()
  assert n_pairs > 0 and n_mapped > 0
  fg = cells > 0
  assert np.array_equal(out > 0, fg)
  pair = np.unique(np.stack([out[fg], cells[fg].astype(np.int64)], axis=1), axis=0)
  assert len(np.unique(pair[:, 0])) == len(pair), 'false merge across a slab face'
  assert len(np.unique(out[fg])) < total          # ids were joined
