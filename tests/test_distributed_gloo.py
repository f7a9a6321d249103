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
