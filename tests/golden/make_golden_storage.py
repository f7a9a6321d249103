"""Output-format fixture: seg-*.npz files written by the REAL reference `ffn.inference.storage.save_subvolume`
(ffn/inference/storage.py:154-171) and what the reference's `load_segmentation` (:414-488) reads back from a file
written by THIS repository's writer.

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden_storage.py

Harness as in make_golden.py (reference modules unmodified, third-party imports stubbed); TensorFlow's gfile is
replaced by os / open.  Outputs (small on purpose):
  ref_subvolume_u8.npz / ref_subvolume_u16.npz   the reference's own files (ids <= 255 / > 255: reduce_id_bits)
  storage_roundtrip.npz                          inputs + `load_segmentation(split_cc=False)` of OUR file, as returned by
                                                 the reference reader
"""
import contextlib
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def inputs(max_id):
  rng = np.random.RandomState(max_id)
  labels = np.zeros((20, 24, 28), dtype=np.int32)
  ids = [3, 17, max_id]
  for k, sid in enumerate(ids):
    labels[2 + 5 * k:7 + 5 * k, 3 + 4 * k:15 + 4 * k, 5:20] = sid
  labels[rng.randint(0, 20, 30), rng.randint(0, 24, 30), rng.randint(0, 28, 30)] = 0
  origins = {sid: ((2 + 5 * k, 4 + 4 * k, 6), 10 + k, 0.25 * (k + 1)) for k, sid in enumerate(ids)}
  overlaps = {sid: np.array([[1, 2], [k, 3 * k]], dtype=np.int64) for k, sid in enumerate(ids)}
  return labels, origins, overlaps


def main():
  os.environ.setdefault('PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION', 'python')
  mg.install_stubs()
  sys.path.insert(0, mg.REF)
  from ffn.inference import storage as ref_storage

  class _GFileCtx:
    def __init__(self, path, mode='r'):
      self._f = open(path, mode)

    def __enter__(self):
      return self._f

    def __exit__(self, *a):
      self._f.close()

  class _Gfile:
    makedirs = staticmethod(lambda p: os.makedirs(p, exist_ok=True))
    exists = staticmethod(os.path.exists)
    GFile = _GFileCtx
  ref_storage.gfile = _Gfile

  @contextlib.contextmanager
  def atomic_file(path, mode='w+b'):
    with open(path, mode) as f:
      yield f
  ref_storage.atomic_file = atomic_file

  tmp = tempfile.mkdtemp(prefix='storage_golden_')
  saved = {}
  for tag, max_id in (('u8', 200), ('u16', 300)):
    labels, origins, overlaps = inputs(max_id)
    ref_origins = {sid: ref_storage.OriginInfo(*o) for sid, o in origins.items()}
    path = os.path.join(tmp, 'ref_%s.npz' % tag)
    ref_storage.save_subvolume(labels, ref_origins, path, request=b'request-bytes', counters='{"a": 1}', overlaps=overlaps)
    shutil.copy(path, os.path.join(HERE, 'ref_subvolume_%s.npz' % tag))
    # the other direction: OUR writer, the reference's reader
    sys.path.insert(0, mg.REPO)
    from ffn_b200.inference import storage as our_storage
    corner = (0, 0, 0)
    our_dir = os.path.join(tmp, 'ours_' + tag)
    our_path = our_storage.segmentation_path(our_dir, corner)
    our_storage.save_subvolume(labels.copy(), {sid: our_storage.OriginInfo(*o) for sid, o in origins.items()}, our_path,
                               request=b'request-bytes', counters='{"a": 1}', overlaps=overlaps)
    assert ref_storage.segmentation_path(our_dir, corner) == our_path            # same path layout
    seg, got_origins = ref_storage.load_segmentation(our_dir, corner, split_cc=False, min_size=0)
    assert seg.dtype == np.uint64 and np.array_equal(seg, labels.astype(np.uint64))
    assert {k: tuple(v) for k, v in got_origins.items()} == {k: tuple(ref_origins[k]) for k in ref_origins}
    saved[tag + '_labels'] = labels
    saved[tag + '_read_by_reference'] = seg
    saved[tag + '_origin_ids'] = np.array(sorted(origins))
    saved[tag + '_origin_start'] = np.array([origins[k][0] for k in sorted(origins)])
    saved[tag + '_origin_iters'] = np.array([origins[k][1] for k in sorted(origins)])
    saved[tag + '_origin_wall'] = np.array([origins[k][2] for k in sorted(origins)])
    print(tag, 'reference file dtype', np.load(path, allow_pickle=True)['segmentation'].dtype, '; our file read by the reference: ok')
  np.savez_compressed(os.path.join(HERE, 'storage_roundtrip.npz'), **saved)
  print('wrote ref_subvolume_u8.npz, ref_subvolume_u16.npz, storage_roundtrip.npz')


if __name__ == '__main__':
  main()
