"""Resegmentation fixture: the REAL reference `ffn.inference.resegmentation.process_point`
(ffn/inference/resegmentation.py:111-293) on the 64x72x80 golden volume, for a PAIR point and an ENDPOINT.

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden_reseg.py

Same harness as make_golden.py: the reference's unmodified modules with stubbed third-party imports, the fp32 oracle
network as the executor client.  What is stubbed around `process_point` is I/O and plumbing only:
  * `runner`: an object with `.counters`, `.init_seg_volume` and `make_canvas(corner, size, keep_history=True)` that does
    what Runner.make_canvas does for a request without masks / alignment (runner.py:307-414): crop, normalise,
    `inference.Canvas(...)`, `init_segmentation_from_volume`;
  * `gfile` / `storage.atomic_file` (TensorFlow file API) -> os / open;
  * `connectomics.segmentation.labels.make_contiguous` (un-vendored) -> its published definition, restated;
  * the module's `np`: a proxy that builds the ragged object arrays `np.array(histories)` / `start_points=` relied on
    before numpy 1.24.
Input segmentation = the reference's own segment_all result on that volume (flood_fill_64.npz).
Output: reseg_64.npz — the request parameters and, per point, everything process_point saved.
"""
import contextlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

RADIUS_ZYX = (24, 28, 32)


def find_points(seg):
  """A pair point (label 0, different objects within 3 voxels on its -x / +x side) and the same point as an endpoint;
  the radius box around it must fit into the volume with one voxel to spare (resegmentation.py:96-106)."""
  rz, ry, rx = RADIUS_ZYX
  best = None
  for z in range(rz, seg.shape[0] - rz - 1):
    for y in range(ry, seg.shape[1] - ry - 1):
      for x in range(rx, seg.shape[2] - rx - 1):
        if seg[z, y, x] != 0:
          continue
        left = [int(v) for v in seg[z, y, max(x - 3, 0):x][::-1] if v > 0]
        right = [int(v) for v in seg[z, y, x + 1:x + 4] if v > 0]
        if left and right and left[0] != right[0]:
          sub = seg[z - rz:z + rz + 1, y - ry:y + ry + 1, x - rx:x + rx + 1]
          score = min(int((sub == left[0]).sum()), int((sub == right[0]).sum()))
          if best is None or score > best[0]:
            best = (score, (z, y, x), left[0], right[0])
  if best is None:
    raise SystemExit('no decision point found')
  return best[1], best[2], best[3]


def main():
  os.environ.setdefault('PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION', 'python')
  mg.install_stubs()
  sys.path.insert(0, mg.REF)
  from ffn.inference import align as ref_align
  from ffn.inference import inference as ref_inference
  from ffn.inference import inference_pb2 as ref_pb2
  from ffn.inference import inference_utils as ref_utils
  from ffn.inference import movement as ref_movement
  from ffn.inference import resegmentation as ref_reseg
  from ffn.inference import storage as ref_storage
  from ffn.training import model as ref_model
  from ffn_b200 import tf_checkpoint
  from oracle.network import ConvStackOracle

  w, b = tf_checkpoint.load_convstack_npz(os.path.join(HERE, 'fib25_convstack.npz'))
  net32 = ConvStackOracle(w, b)

  class Client:
    def start(self):
      return 0

    def finish(self):
      pass

    def predict(self, seed, image, fetches):
      return {'logits': net32(seed, image)[..., np.newaxis]}

  # ---- I/O shims (TensorFlow's gfile is a stub here)
  class _Gfile:
    makedirs = staticmethod(lambda p: os.makedirs(p, exist_ok=True))
    exists = staticmethod(os.path.exists)
  ref_reseg.gfile = _Gfile

  @contextlib.contextmanager
  def atomic_file(path, mode='w+b'):
    with open(path, mode) as f:
      yield f
  ref_storage.atomic_file = atomic_file

  # connectomics.segmentation.labels is an un-vendored dependency (stubbed): restate its published make_contiguous —
  # ids in ascending order (0 stays 0) -> 0..n-1, returning the relabelled array and the (global, local) pairs
  def make_contiguous(labels):
    orig = np.unique(np.append(labels, np.uint64(0)))
    new = np.arange(len(orig), dtype=np.uint64)
    return new[np.searchsorted(orig, labels)], list(zip(orig.tolist(), new.tolist()))
  ref_inference.label_utils.make_contiguous = make_contiguous

  # process_point's final np.array(histories) / savez(start_points=[[...], [...]]) are ragged: numpy >= 1.24 refuses
  # the implicit object array the code was written for.  The module sees numpy through a proxy that restores it.
  class _Numpy:
    def __getattr__(self, name):
      return getattr(np, name)

    @staticmethod
    def _ragged(items):
      out = np.empty(len(items), dtype=object)
      for i, item in enumerate(items):
        out[i] = item
      return out

    def array(self, obj, *a, **k):
      try:
        return np.array(obj, *a, **k)
      except ValueError:
        return self._ragged(obj)

    def savez_compressed(self, fd, **items):
      fixed = {}
      for key, val in items.items():
        try:
          fixed[key] = np.asanyarray(val)
        except ValueError:
          fixed[key] = self._ragged([np.asarray(v) for v in val])
      return np.savez_compressed(fd, **fixed)
  ref_reseg.np = _Numpy()

  g = np.load(os.path.join(HERE, 'flood_fill_64.npz'))
  volume = g['volume']
  seg = np.maximum(g['segmentation'], 0).astype(np.uint64)
  point, id_a, id_b = find_points(seg)

  req = ref_pb2.ResegmentationRequest()
  opts = req.inference.inference_options
  opts.init_activation = 0.95
  opts.pad_value = 0.05
  opts.move_threshold = 0.9
  opts.segment_threshold = 0.6
  opts.min_segment_size = 1000
  opts.min_boundary_dist.x = opts.min_boundary_dist.y = opts.min_boundary_dist.z = 1
  req.inference.image_mean = 128
  req.inference.image_stddev = 33
  req.radius.z, req.radius.y, req.radius.x = RADIUS_ZYX
  req.max_retry_iters = 3
  req.exclusion_radius.x = req.exclusion_radius.y = req.exclusion_radius.z = 3
  req.init_exclusion_radius.x = req.init_exclusion_radius.y = req.init_exclusion_radius.z = 2
  req.analysis_radius.z, req.analysis_radius.y, req.analysis_radius.x = 8, 12, 16
  req.segment_recovery_fraction = 0.4
  out_dir = tempfile.mkdtemp(prefix='reseg_golden_')
  req.output_directory = out_dir
  for ids in ((id_a, id_b), (id_a,)):
    pt = req.points.add()
    pt.id_a = ids[0]
    if len(ids) > 1:
      pt.id_b = ids[1]
    pt.point.z, pt.point.y, pt.point.x = point

  info = ref_model.ModelInfo(np.array([8, 8, 8]), np.array([33, 33, 33]), np.array([33, 33, 33]), np.array([33, 33, 33]))

  class StubRunner:
    def __init__(self):
      self.counters = ref_utils.Counters()
      self.init_seg_volume = seg[np.newaxis]

    def make_canvas(self, corner, subvol_size, **canvas_kwargs):
      corner, subvol_size = np.asarray(corner), np.asarray(subvol_size)
      alignment = ref_align.Aligner().generate_alignment(corner, subvol_size)
      end = corner + subvol_size
      image = volume[corner[0]:end[0], corner[1]:end[1], corner[2]:end[2]]
      image = (image.astype(np.float32) - req.inference.image_mean) / req.inference.image_stddev
      canvas = ref_inference.Canvas(
          info, Client(), image, req.inference.inference_options, counters=self.counters.get_sub_counters(),
          movement_policy_fn=ref_movement.get_policy_fn(req.inference, info), corner_zyx=corner,
          keep_probability_maps=True, **canvas_kwargs)
      canvas.init_segmentation_from_volume(self.init_seg_volume, corner, end, None)
      return canvas, alignment

  runner = StubRunner()
  saved = {}
  for n, ids in enumerate(((id_a, id_b), (id_a,))):
    ref_reseg.process_point(req, runner, n, voxel_size=(1, 1, 1))
    name = '%d-%d_at_%d_%d_%d.npz' % (ids[0], ids[1] if len(ids) > 1 else 0, point[2], point[1], point[0])
    out = np.load(os.path.join(out_dir, name), allow_pickle=True)
    tag = 'pair' if len(ids) > 1 else 'endpoint'
    saved[tag + '_raw_probs'] = out['raw_probs']
    saved[tag + '_probs'] = out['probs']
    saved[tag + '_corner_zyx'] = np.asarray(out['corner_zyx'])
    saved[tag + '_is_shift'] = np.asarray(bool(out['is_shift']))
    hist, dele, starts = out['histories'], out['deletes'], out['start_points']
    saved[tag + '_n_objects'] = np.asarray(len(hist))
    for k in range(len(hist)):
      saved['%s_history_%d' % (tag, k)] = np.asarray(hist[k], dtype=np.int32).reshape(-1, 3)
      saved['%s_deletes_%d' % (tag, k)] = np.asarray(dele[k], dtype=np.int64)
    for k in range(2):
      saved['%s_starts_%d' % (tag, k)] = np.asarray(starts[k], dtype=np.int64).reshape(-1, 3)
    print(tag, 'objects', len(hist), 'steps', [len(h) for h in hist], 'starts', [list(map(tuple, np.asarray(s).reshape(-1, 3))) for s in starts],
          'voxels >= 0.6', [int((p >= 154).sum()) for p in out['raw_probs']])
  np.savez_compressed(os.path.join(HERE, 'reseg_64.npz'), point_zyx=np.asarray(point), id_a=np.asarray(id_a),
                      id_b=np.asarray(id_b), radius_zyx=np.asarray(RADIUS_ZYX), max_retry_iters=np.asarray(3),
                      exclusion_radius=np.asarray(3), init_exclusion_radius=np.asarray(2),
                      analysis_radius_zyx=np.asarray((8, 12, 16)), segment_recovery_fraction=np.asarray(0.4), **saved)
  print('wrote reseg_64.npz; point', point, 'ids', id_a, id_b)


if __name__ == '__main__':
  main()
