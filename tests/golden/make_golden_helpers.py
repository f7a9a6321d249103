"""Small host helpers pinned against the REAL reference functions (unmodified modules, harness as in make_golden.py):

  * align.Alignment: align_and_crop (source partly / fully outside the destination box, fill value), expand_bounds,
    transform, rescaled, transform_shift_mask; Aligner.generate_alignment            (align.py:20-172)
  * storage: clip_subvolume_to_bounds on 3-d / 4-d volumes (:302-320), dequantize_probability (:146-151),
    threshold_segmentation from a .prob file (:275-288), the path helpers / get_existing_subvolume_path /
    get_existing_corners / get_corner_from_path (:174-272)
  * segmentation.reduce_id_bits and clear_dust                                         (segmentation.py:66-86, clear_dust)

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden_helpers.py  ->  helpers_ref.npz
"""
import contextlib
import glob
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

CROPS = [  # (src_corner, src_shape, dst_corner, dst_size, fill)
    ((2, 3, 4), (6, 7, 8), (2, 3, 4), (6, 7, 8), 0),          # identical boxes
    ((2, 3, 4), (6, 7, 8), (3, 4, 5), (4, 4, 4), 0),          # destination inside the source
    ((2, 3, 4), (6, 7, 8), (0, 1, 2), (5, 6, 7), 9),          # overlap at the low corner, fill 9
    ((2, 3, 4), (6, 7, 8), (6, 8, 10), (5, 5, 5), 0),         # overlap at the high corner
    ((2, 3, 4), (3, 3, 3), (10, 10, 10), (2, 2, 2), 7),       # disjoint
]
CLIPS = [((0, 0, 0), (5, 5, 5)), ((-3, 2, 10), (8, 8, 30)), ((15, 20, 26), (10, 10, 10)), ((4, 4, 4), (3, 2, 1))]
CORNERS = [(0, 0, 0), (64, 128, 256), (7, 0, 33)]


def main():
  os.environ.setdefault('PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION', 'python')
  mg.install_stubs()
  sys.path.insert(0, mg.REF)
  from ffn.inference import align as ref_align
  from ffn.inference import segmentation as ref_seg
  from ffn.inference import storage as ref_storage
  from ffn.utils import bounding_box as ref_bbox

  class _BBox(ref_bbox.BoundingBox):         # see make_golden_build_mask.py
    def intersection(self, other):
      return ref_bbox.intersection(self, other)
  ref_storage.bounding_box = types.SimpleNamespace(BoundingBox=_BBox)

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
    glob = staticmethod(glob.glob)
    GFile = _GFileCtx
  ref_storage.gfile = _Gfile

  class _Numpy:                              # align.py still says np.int (removed in numpy 1.24): the alias is all that is added
    int = int

    def __getattr__(self, name):
      return getattr(np, name)
  ref_align.np = _Numpy()

  out = {}
  rng = np.random.RandomState(5)
  # ---- align
  for i, (sc, ss, dc, ds, fill) in enumerate(CROPS):
    src = rng.randint(1, 100, size=ss).astype(np.int32)
    a = ref_align.Alignment(dc, ds)
    out['crop_src_%d' % i] = src
    out['crop_out_%d' % i] = a.align_and_crop(np.array(sc), src, np.array(dc), np.array(ds), fill=fill)
  a = ref_align.Aligner().generate_alignment((4, 5, 6), (10, 20, 30))
  out['align_corner'], out['align_size'] = np.asarray(a.corner), np.asarray(a.size)
  for fwd in (True, False):
    c, s = a.expand_bounds(np.array((1, 2, 3)), np.array((7, 8, 9)), forward=fwd)
    out['expand_%d' % fwd] = np.stack([np.asarray(c), np.asarray(s)])
  pts = np.array([[1, 2, 3], [4, 5, 6]]).T
  out['transform'] = np.asarray(a.transform(pts))
  r = a.rescaled(np.array((1.0, 0.5, 0.5)))
  out['rescaled'] = np.stack([np.asarray(r.corner), np.asarray(r.size)]).astype(np.float64)
  shift = rng.rand(2, 4, 5, 6).astype(np.float32)
  out['shift_in'] = shift
  out['shift_out'] = np.asarray(a.transform_shift_mask(np.array((4, 5, 6)), 2, shift))
  # ---- storage: clip
  vol3, vol4 = np.zeros((20, 24, 28)), np.zeros((2, 20, 24, 28))
  clips = []
  for corner, size in CLIPS:
    for vol in (vol3, vol4):
      try:
        c, s = ref_storage.clip_subvolume_to_bounds(np.array(corner), np.array(size), vol)
        clips.append(list(np.asarray(c).astype(int)) + list(np.asarray(s).astype(int)))
      except Exception:  # disjoint: the vendored intersection() returns None
        clips.append([-1] * 6)
  out['clips'] = np.asarray(clips)
  # ---- storage: probabilities
  q = np.arange(256, dtype=np.uint8)
  out['dequantized'] = ref_storage.dequantize_probability(q)
  tmp = tempfile.mkdtemp(prefix='helpers_golden_')
  corner = (5, 6, 7)
  labels = rng.randint(0, 4, size=(6, 7, 8)).astype(np.uint64)
  qprob = rng.randint(0, 256, size=(6, 7, 8)).astype(np.uint8)
  prob_path = ref_storage.object_prob_path(tmp, corner)
  os.makedirs(os.path.dirname(prob_path), exist_ok=True)
  np.savez_compressed(prob_path.replace('.prob', '.prob.npz'), qprob=qprob)
  os.rename(prob_path.replace('.prob', '.prob.npz'), prob_path)
  thr = labels.copy()
  ref_storage.threshold_segmentation(tmp, corner, thr, 0.7)
  out['thr_labels'], out['thr_qprob'], out['thr_out'] = labels, qprob, thr
  # ---- storage: paths
  paths = {}
  for c in CORNERS:
    paths[str(c)] = [ref_storage.subvolume_path('/out', c, 'npz'), ref_storage.legacy_subvolume_path('/out', c, 'npz'),
                     ref_storage.segmentation_path('/out', c), ref_storage.object_prob_path('/out', c),
                     ref_storage.checkpoint_path('/out', c), ref_storage.legacy_segmentation_path('/out', c),
                     ref_storage.legacy_object_prob_path('/out', c)]
    assert ref_storage.get_corner_from_path(paths[str(c)][0]) == tuple(c)
  d2 = tempfile.mkdtemp(prefix='helpers_golden_')
  for c, legacy, suffix in (((1, 2, 3), False, 'npz'), ((4, 5, 6), True, 'npz'), ((7, 8, 9), False, 'cpoint')):
    p = ref_storage.legacy_subvolume_path(d2, c, suffix) if legacy else ref_storage.subvolume_path(d2, c, suffix)
    os.makedirs(os.path.dirname(p), exist_ok=True)
    open(p, 'wb').close()
  existing = {}
  for c in ((1, 2, 3), (4, 5, 6), (7, 8, 9), (9, 9, 9)):
    for allow in (False, True):
      got = ref_storage.get_existing_subvolume_path(d2, c, allow)
      existing['%r/%d' % (c, allow)] = None if got is None else os.path.relpath(got, d2)
  out['paths_json'] = np.asarray(json.dumps(paths))
  out['existing_json'] = np.asarray(json.dumps(existing))
  out['existing_corners'] = np.asarray(sorted(ref_storage.get_existing_corners(d2)))
  # ---- segmentation helpers
  for top in (200, 300, 70000):
    lab = rng.randint(0, top + 1, size=(4, 5, 6)).astype(np.int64)
    lab[0, 0, 0] = top
    out['reduce_in_%d' % top] = lab
    out['reduce_dtype_%d' % top] = np.asarray(str(ref_seg.reduce_id_bits(lab).dtype))
  dust = np.zeros((6, 8, 8), dtype=np.uint64)
  dust[0:2, 0:2, 0:2] = 5
  dust[3:6, 2:7, 1:8] = 9
  dust[0, 7, 7] = 11
  out['dust_in'] = dust
  cleaned = ref_seg.clear_dust(dust.copy(), min_size=9)
  out['dust_out'] = cleaned
  np.savez_compressed(os.path.join(HERE, 'helpers_ref.npz'), **out)
  print('wrote helpers_ref.npz;', 'clips', out['clips'].tolist()[:4], 'existing', existing)


if __name__ == '__main__':
  main()
