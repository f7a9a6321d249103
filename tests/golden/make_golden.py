"""Generates the golden fixtures in tests/golden/ by running the REAL reference Python modules.

Run once in the build container (where /root/reference exists):

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden.py

The reference's flood-fill logic (ffn/inference/{inference,movement,seed,storage}.py) is pure
numpy/scipy, but every module imports TensorFlow / connectomics / skimage / edt / jax / h5py at
module scope and none of those are installable offline.  This script registers inert stub modules
for those imports, imports the unmodified reference modules from /root/reference, and drives
`Canvas.segment_at` / `Canvas.segment_all` / `movement.get_scored_move_offsets` /
`storage.quantize_probability` / `seed.PolicyGrid3d` with an executor client whose `predict`
evaluates the network with oracle/network.py (torch CPU conv3d on the shipped FIB-25 checkpoint;
TensorFlow's Conv3D itself cannot run here, see oracle/__init__.py).

Outputs (small, committed):
  flood_fill_64.npz   - trajectory, seed/segmentation/seg_prob canvases, origins, overlaps and
                        counters of the reference Canvas on a 64x72x80 phantom (grid seeds)
  moves.npz           - get_scored_move_offsets on random logit patches (iso + aniso deltas)
  qprob.npz           - quantize_probability known answers
  net_patches.npz     - a few (seed, image) -> logits patches of the oracle network itself
                        (fp32 and fp64), used by the GPU parity tests on the GPU box
  sample_training2_summary.json - sanity ranges read from results/fib25/sample-training2.npz
"""

import json
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402


class _Stub(types.ModuleType):
  """Module that answers every attribute with another stub (callable, subscriptable)."""

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    child = _Stub(self.__name__ + '.' + name)
    setattr(self, name, child)
    return child

  def __call__(self, *a, **k):
    return _Stub(self.__name__ + '()')

  def __getitem__(self, k):
    return _Stub(self.__name__ + '[]')

  def __mro_entries__(self, bases):
    return (object,)


def install_stubs():
  names = [
      'tensorflow', 'tensorflow.compat', 'tensorflow.compat.v1', 'tensorflow.io',
      'tensorflow.io.gfile', 'tf_slim', 'jax', 'jax.numpy', 'h5py', 'tensorstore', 'edt',
      'skimage', 'skimage.feature', 'skimage.morphology', 'skimage.exposure',
      'connectomics', 'connectomics.common', 'connectomics.common.bounding_box',
      'connectomics.common.import_util', 'connectomics.segmentation',
      'connectomics.segmentation.labels', 'connectomics.common.utils',
      'connectomics.common.box_generator', 'connectomics.volume', 'connectomics.volume.metadata',
  ]
  for n in names:
    m = _Stub(n)
    m.__path__ = []  # mark as package
    sys.modules[n] = m
  for n in names:
    if '.' in n:
      parent, child = n.rsplit('.', 1)
      setattr(sys.modules[parent], child, sys.modules[n])


def main():
  os.environ.setdefault('PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION', 'python')
  install_stubs()
  sys.path.insert(0, REF)
  from ffn.inference import inference as ref_inference
  from ffn.inference import inference_pb2 as ref_pb2
  from ffn.inference import movement as ref_movement
  from ffn.inference import seed as ref_seed
  from ffn.inference import storage as ref_storage
  from ffn.training import model as ref_model

  from ffn_b200 import tf_checkpoint
  from ffn_b200.synthetic import voronoi_phantom
  from oracle.network import ConvStackOracle

  ckpt = os.path.join(REF, 'models/fib25/model.ckpt-27465036')
  w, b = tf_checkpoint.load_convstack_weights(ckpt, 12)
  net32 = ConvStackOracle(w, b)
  # The GPU box has no /root/reference: ship the (data-only) FIB-25 parameters as a fixture.
  np.savez_compressed(os.path.join(OUT, 'fib25_convstack.npz'),
                      **{'w%02d' % i: x for i, x in enumerate(w)},
                      **{'b%02d' % i: x for i, x in enumerate(b)})

  class Client:
    """Stands where ThreadingExecutorClient stands (executor.py:111-139)."""

    def __init__(self):
      self.calls = 0

    def start(self):
      return 0

    def finish(self):
      pass

    def predict(self, seed, image, fetches):
      self.calls += 1
      return {'logits': net32(seed, image)[..., np.newaxis]}

  # ---------------------------------------------------------------- flood fill (segment_all)
  shape = (64, 72, 80)
  vol = voronoi_phantom(shape, seed=11, cell_volume=40000.0)
  image = (vol.astype(np.float32) - 128.0) / 33.0         # runner.py:383-385
  opts = ref_pb2.InferenceOptions()
  opts.init_activation = 0.95
  opts.pad_value = 0.05
  opts.move_threshold = 0.9
  opts.segment_threshold = 0.6
  opts.min_segment_size = 1000
  opts.min_boundary_dist.x = 1
  opts.min_boundary_dist.y = 1
  opts.min_boundary_dist.z = 1
  request = ref_pb2.InferenceRequest()
  request.inference_options.CopyFrom(opts)
  info = ref_model.ModelInfo(np.array([8, 8, 8]), np.array([33, 33, 33]),
                             np.array([33, 33, 33]), np.array([33, 33, 33]))
  client = Client()
  canvas = ref_inference.Canvas(
      info, client, image, opts,
      movement_policy_fn=ref_movement.get_policy_fn(request, info),
      keep_probability_maps=True, keep_history=True)

  trace = []
  seg_starts = []
  orig_update = canvas.update_at

  def traced_update(pos):
    trace.append(tuple(int(p) for p in pos))
    return orig_update(pos)
  canvas.update_at = traced_update
  orig_segment_at = canvas.segment_at

  def traced_segment_at(pos, **kw):
    seg_starts.append((tuple(int(p) for p in pos), len(trace)))
    return orig_segment_at(pos, **kw)
  canvas.segment_at = traced_segment_at

  canvas.segment_all(seed_policy=ref_seed.PolicyGrid3d)
  seeds = np.asarray(canvas.seed_policy.coords, dtype=np.int64)

  origins = np.array([(k,) + tuple(v.start_zyx) + (v.iters,) for k, v in sorted(canvas.origins.items())],
                     dtype=np.int64).reshape(-1, 5)
  ov_ids, ov_cnt, ov_owner = [], [], []
  for k, v in sorted(canvas.overlaps.items()):
    for i, c in zip(v[0], v[1]):
      ov_owner.append(k); ov_ids.append(int(i)); ov_cnt.append(int(c))
  counters = {k: c.value for k, c in canvas.counters if not k.endswith('-time-ms')}
  np.savez_compressed(
      os.path.join(OUT, 'flood_fill_64.npz'),
      volume=vol, seeds=seeds, trace=np.asarray(trace, dtype=np.int32),
      segment_starts=np.asarray([s[0] + (s[1],) for s in seg_starts], dtype=np.int64),
      seed_canvas=np.asarray(canvas.seed), segmentation=np.asarray(canvas.segmentation),
      seg_prob=np.asarray(canvas.seg_prob), origins=origins,
      overlaps=np.asarray([ov_owner, ov_ids, ov_cnt], dtype=np.int64),
      counters=json.dumps(counters),
      thresholds=np.asarray([canvas.options.init_activation, canvas.options.pad_value,
                             canvas.options.move_threshold, canvas.options.segment_threshold],
                            dtype=np.float64),
      policy_threshold=np.float64(canvas.movement_policy.score_threshold))
  print('flood fill: %d steps, %d segment_at calls, %d segments, counters=%s' %
        (len(trace), len(seg_starts), len(canvas.origins), counters))

  # ---------------------------------------------------------------- toy-net flood fill (logic pin)
  from oracle.toy_net import toy_net, toy_image
  tshape = (56, 72, 88)
  _, tcells = voronoi_phantom(tshape, seed=3, cell_volume=30000.0, return_cells=True)
  timage = toy_image(tcells)

  class ToyClient(Client):
    def predict(self, seed, image, fetches):
      self.calls += 1
      return {'logits': toy_net(seed, image)[..., np.newaxis]}

  topts = ref_pb2.InferenceOptions()
  topts.CopyFrom(opts)
  topts.min_segment_size = 3000
  topts.min_boundary_dist.x = 2
  topts.min_boundary_dist.y = 1
  topts.min_boundary_dist.z = 3
  treq = ref_pb2.InferenceRequest()
  treq.inference_options.CopyFrom(topts)
  tcanvas = ref_inference.Canvas(info, ToyClient(), timage, topts,
                                 movement_policy_fn=ref_movement.get_policy_fn(treq, info),
                                 keep_probability_maps=True)
  ttrace = []
  tou = tcanvas.update_at
  tcanvas.update_at = lambda pos: (ttrace.append(tuple(int(p) for p in pos)), tou(pos))[1]
  tcanvas.segment_all(seed_policy=ref_seed.PolicyGrid3d)
  torigins = np.array([(k,) + tuple(v.start_zyx) + (v.iters,)
                       for k, v in sorted(tcanvas.origins.items())], dtype=np.int64).reshape(-1, 5)
  to_owner, to_ids, to_cnt = [], [], []
  for k, v in sorted(tcanvas.overlaps.items()):
    for i, c in zip(v[0], v[1]):
      to_owner.append(k); to_ids.append(int(i)); to_cnt.append(int(c))
  tcounters = {k: c.value for k, c in tcanvas.counters if not k.endswith('-time-ms')}
  np.savez_compressed(
      os.path.join(OUT, 'toy_flood_fill.npz'), cells=tcells,
      seeds=np.asarray(tcanvas.seed_policy.coords, dtype=np.int64),
      trace=np.asarray(ttrace, dtype=np.int32), seed_canvas=np.asarray(tcanvas.seed),
      segmentation=np.asarray(tcanvas.segmentation), seg_prob=np.asarray(tcanvas.seg_prob),
      origins=torigins, overlaps=np.asarray([to_owner, to_ids, to_cnt], dtype=np.int64),
      counters=json.dumps(tcounters), min_segment_size=3000, min_boundary_dist=np.asarray([3, 1, 2]))
  print('toy flood fill: %d steps, %d segments, counters=%s' % (len(ttrace), len(tcanvas.origins), tcounters))

  # ---------------------------------------------------------------- single segment_at
  client2 = Client()
  c2 = ref_inference.Canvas(info, client2, image, opts,
                            movement_policy_fn=ref_movement.get_policy_fn(request, info))
  t2 = []
  ou = c2.update_at
  c2.update_at = lambda pos: (t2.append(tuple(int(p) for p in pos)), ou(pos))[1]
  start = tuple(int(v) for v in origins[int(np.argmax(origins[:, 4])), 1:4])
  iters = c2.segment_at(start)
  np.savez_compressed(os.path.join(OUT, 'segment_at_64.npz'), start=np.asarray(start),
                      iters=np.int64(iters), trace=np.asarray(t2, dtype=np.int32),
                      seed_canvas=np.asarray(c2.seed),
                      queue=np.asarray([(float(s),) + tuple(int(v) for v in p)
                                        for s, p in c2.movement_policy.scored_coords],
                                       dtype=np.float64).reshape(-1, 4))
  print('segment_at: start %s iters %d' % (start, iters))

  # ---------------------------------------------------------------- movement known answers
  rng = np.random.RandomState(5)
  th = float(canvas.movement_policy.score_threshold)
  cases = []
  for deltas, shp in (((8, 8, 8), (33, 33, 33)), ((4, 8, 8), (17, 33, 33)), ((0, 8, 8), (1, 33, 33))):
    for k in range(12):
      lg = (rng.randn(*shp) * 2.0 + (k % 3)).astype(np.float32)
      if k == 5:   # exact ties on a face
        lg[...] = np.float32(3.0)
      if k == 6:   # nothing crosses the threshold
        lg[...] = np.float32(-5.0)
      moves = sorted(ref_movement.get_scored_move_offsets(np.array(deltas), lg, threshold=th),
                     reverse=True)
      arr = np.asarray([(float(s),) + tuple(int(v) for v in r) for s, r in moves],
                       dtype=np.float64).reshape(-1, 4)
      cases.append((np.asarray(deltas), lg, arr))
  np.savez_compressed(os.path.join(OUT, 'moves.npz'), n=len(cases), threshold=th,
                      **{'deltas_%d' % i: c[0] for i, c in enumerate(cases)},
                      **{'logits_%d' % i: c[1] for i, c in enumerate(cases)},
                      **{'moves_%d' % i: c[2] for i, c in enumerate(cases)})

  # ---------------------------------------------------------------- qprob known answers
  p = np.concatenate([np.array([0, .003, .5, .6, .999, 1.0, np.nan]), rng.rand(2000),
                      np.arange(255) / 254.0]).astype(np.float32)
  np.savez_compressed(os.path.join(OUT, 'qprob.npz'), prob=p, q=ref_storage.quantize_probability(p))

  # ---------------------------------------------------------------- network patches (oracle itself)
  net64 = ConvStackOracle(w, b, dtype=__import__('torch').float64)
  seeds_p, imgs_p, l32, l64 = [], [], [], []
  for k, pos in enumerate(trace[:40:8]):
    sel = tuple(slice(p - 16, p + 17) for p in pos)
    s = np.full((33, 33, 33), np.float32(canvas.options.pad_value), np.float32)
    if k:
      s = np.where(np.isnan(canvas.seed[sel]), np.float32(canvas.options.pad_value),
                   canvas.seed[sel]).astype(np.float32)
    else:
      s[16, 16, 16] = np.float32(canvas.options.init_activation)
    seeds_p.append(s); imgs_p.append(image[sel])
    l32.append(net32(s, image[sel])); l64.append(net64(s, image[sel]))
  np.savez_compressed(os.path.join(OUT, 'net_patches.npz'), seed=np.asarray(seeds_p),
                      image=np.asarray(imgs_p), logits_fp32=np.asarray(l32),
                      logits_fp64=np.asarray(l64))

  # ---------------------------------------------------------------- shipped golden result summary
  gpath = os.path.join(REF, 'results/fib25/sample-training2.npz')
  mod = types.ModuleType('google3.research.neuromancer.segmentation.ffn.storage')
  import collections
  mod.OriginInfo = collections.namedtuple('OriginInfo', ['start_zyx', 'iters', 'walltime_sec'])
  for i, part in enumerate('google3.research.neuromancer.segmentation.ffn.storage'.split('.')):
    full = '.'.join('google3.research.neuromancer.segmentation.ffn.storage'.split('.')[:i + 1])
    sys.modules.setdefault(full, types.ModuleType(full))
  sys.modules['google3.research.neuromancer.segmentation.ffn.storage'] = mod
  g = np.load(gpath, allow_pickle=True, encoding='latin1')
  seg = g['segmentation']
  org = g['origins'].item()
  raw = g['counters'].item()
  cnt = json.loads(raw.decode('utf-8') if isinstance(raw, bytes) else str(raw))
  summary = {
      'shape': list(seg.shape), 'dtype': str(seg.dtype), 'num_segments': int(len(np.unique(seg)) - 1),
      'filled_fraction': float((seg > 0).mean()), 'num_origins': len(org),
      'origins_carry_own_id': int(sum(int(seg[tuple(v.start_zyx)]) == int(k) for k, v in org.items())),
      'total_iters_in_origins': int(sum(v.iters for v in org.values())),
      'counters': {k: v for k, v in cnt.items()},
  }
  with open(os.path.join(OUT, 'sample_training2_summary.json'), 'w') as f:
    json.dump(summary, f, indent=1, sort_keys=True)
  print('golden summary:', {k: summary[k] for k in ('shape', 'num_segments', 'filled_fraction')})


if __name__ == '__main__':
  main()
