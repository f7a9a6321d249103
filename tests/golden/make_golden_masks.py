"""Fixtures pinning the masked and the anisotropic flood fill, from the REAL reference modules.

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden_masks.py

Same harness as make_golden.py, machine-independent toy network (oracle/toy_net.py):
  toy_masks_flood_fill.npz  Canvas.segment_all with a MovementRestrictor carrying `mask`, `seed_mask` and a
                            shift mask (movement.py:247-336; inference.py:507-509, :562-568)
  toy_aniso_flood_fill.npz  fov (z, y, x) = (17, 33, 33), deltas (4, 8, 8): the anisotropic geometry of
                            BASELINE configs[4] (movement.py:42-100 face sizes 17x17 / 9x17)
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def record(canvas, trace):
  origins = np.array([(k,) + tuple(v.start_zyx) + (v.iters,) for k, v in sorted(canvas.origins.items())],
                     dtype=np.int64).reshape(-1, 5)
  owner, ids, cnt = [], [], []
  for k, v in sorted(canvas.overlaps.items()):
    for i, c in zip(v[0], v[1]):
      owner.append(k); ids.append(int(i)); cnt.append(int(c))
  counters = {k: c.value for k, c in canvas.counters if not k.endswith('-time-ms')}
  return dict(seeds=np.asarray(canvas.seed_policy.coords, dtype=np.int64), trace=np.asarray(trace, dtype=np.int32).reshape(-1, 3),
              seed_canvas=np.asarray(canvas.seed), segmentation=np.asarray(canvas.segmentation),
              seg_prob=np.asarray(canvas.seg_prob), origins=origins,
              overlaps=np.asarray([owner, ids, cnt], dtype=np.int64), counters=json.dumps(counters))


def main():
  os.environ.setdefault('PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION', 'python')
  mg.install_stubs()
  sys.path.insert(0, mg.REF)
  from ffn.inference import inference as ref_inference
  from ffn.inference import inference_pb2 as ref_pb2
  from ffn.inference import movement as ref_movement
  from ffn.inference import seed as ref_seed
  from ffn.training import model as ref_model
  from ffn_b200.synthetic import voronoi_phantom
  from oracle.toy_net import toy_net, toy_image

  class ToyClient:
    def start(self):
      return 0

    def finish(self):
      pass

    def predict(self, seed, image, fetches):
      return {'logits': toy_net(seed, image)[..., np.newaxis]}

  opts = ref_pb2.InferenceOptions()
  opts.init_activation, opts.pad_value, opts.move_threshold, opts.segment_threshold = 0.95, 0.05, 0.9, 0.6
  opts.min_segment_size = 2000
  opts.min_boundary_dist.x, opts.min_boundary_dist.y, opts.min_boundary_dist.z = 1, 2, 1
  req = ref_pb2.InferenceRequest()
  req.inference_options.CopyFrom(opts)

  def run(shape, info, restrictor, cells_seed):
    _, cells = voronoi_phantom(shape, seed=cells_seed, cell_volume=30000.0, return_cells=True)
    image = toy_image(cells)
    canvas = ref_inference.Canvas(info, ToyClient(), image, opts, restrictor=restrictor,
                                  movement_policy_fn=ref_movement.get_policy_fn(req, info), keep_probability_maps=True)
    trace = []
    upd = canvas.update_at
    canvas.update_at = lambda pos: (trace.append(tuple(int(p) for p in pos)), upd(pos))[1]
    canvas.segment_all(seed_policy=ref_seed.PolicyGrid3d)
    out = record(canvas, trace)
    out['cells'] = cells
    return out

  # ---- masks
  shape = (56, 72, 88)
  rng = np.random.RandomState(23)
  mask = np.zeros(shape, dtype=bool)
  mask[:, 30:38, :] = True                      # a slab the FoV may not enter
  mask |= rng.rand(*shape) > 0.97
  seed_mask = rng.rand(*shape) > 0.5            # half of the grid seeds are not used
  shift = np.zeros((2,) + (shape[0], shape[1] // 2, shape[2] // 2), dtype=np.float32)
  shift[0, 20:24, 5:8, 30:34] = 6.0             # a distorted patch in four sections
  fov = types.SimpleNamespace(start=np.array([-8, -8, -2]), end=np.array([-8, -8, -2]) + np.array([17, 17, 5]))
  restrictor = ref_movement.MovementRestrictor(mask=mask, seed_mask=seed_mask, shift_mask=shift, shift_mask_fov=fov,
                                               shift_mask_threshold=4, shift_mask_scale=2)
  info = ref_model.ModelInfo(np.array([8, 8, 8]), np.array([33, 33, 33]), np.array([33, 33, 33]), np.array([33, 33, 33]))
  out = run(shape, info, restrictor, 5)
  np.savez_compressed(os.path.join(HERE, 'toy_masks_flood_fill.npz'), mask=mask, seed_mask=seed_mask, shift=shift,
                      fov_start=fov.start, fov_size=np.array([17, 17, 5]), shift_scale=2, shift_threshold=4,
                      min_segment_size=2000, min_boundary_dist=np.asarray([1, 2, 1]), **out)
  print('masks: %d steps, %d segments, counters=%s' % (len(out['trace']), len(out['origins']), out['counters']))

  # ---- anisotropic geometry (xyz: fov 33, 33, 17; deltas 8, 8, 4)
  ainfo = ref_model.ModelInfo(np.array([8, 8, 4]), np.array([33, 33, 17]), np.array([33, 33, 17]), np.array([33, 33, 17]))
  out = run((40, 72, 80), ainfo, None, 9)
  np.savez_compressed(os.path.join(HERE, 'toy_aniso_flood_fill.npz'), min_segment_size=2000,
                      min_boundary_dist=np.asarray([1, 2, 1]), **out)
  print('aniso: %d steps, %d segments, counters=%s' % (len(out['trace']), len(out['origins']), out['counters']))


if __name__ == '__main__':
  main()
