"""Seed-policy and counter fixtures from the REAL reference modules (unmodified; harness as in make_golden.py):

  * `ffn.inference.seed`: PolicyMax, PolicyGrid2d, PolicyGrid3d (custom step / offsets), PolicyDenseSeeds (threshold,
    erosions, invert), ReverseCoords, SequentialPolicies — each iterated to exhaustion through `BaseSeedPolicy.__next__`
    (border filter) on one small canvas (seed.py:36-95,307-313,411-544); the un-vendored `skimage.morphology.
    binary_erosion` PolicyDenseSeeds calls is injected by definition;
  * `ffn.inference.inference_utils.Counters`: `dumps()` of a populated counter tree and the state after `loads()`
    (inference_utils.py:90-150) — the format of counters.txt and of the `counters` item of seg-*.npz / checkpoints.

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden_policies.py  ->  seed_policies_ref.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

SHAPE = (12, 20, 24)
MARGIN = (2, 3, 4)
CASES = [
    ('max', 'PolicyMax', {}),
    ('grid2d', 'PolicyGrid2d', {'step': 6, 'offsets': (0, 3, 1)}),
    ('grid3d', 'PolicyGrid3d', {'step': 5, 'offsets': (0, 2, 4)}),
    ('dense', 'PolicyDenseSeeds', {'threshold': 0.5, 'num_erosions': 1}),
    ('dense_inv', 'PolicyDenseSeeds', {'threshold': 0.55, 'num_erosions': 2, 'invert': True}),
    ('reverse', 'ReverseCoords', {'policy_to_reverse': 'PolicyGrid3d', 'step': 6}),
    ('sequential', 'SequentialPolicies', {'policies': [('PolicyGrid3d', {'step': 7}), ('PolicyGrid2d', {'step': 9})]}),
]


def make_image():
  rng = np.random.RandomState(9)
  from scipy import ndimage
  return ndimage.gaussian_filter(rng.rand(*SHAPE), 1.5).astype(np.float32) * 4 - 1.5


def main():
  os.environ.setdefault('PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION', 'python')
  mg.install_stubs()
  sys.path.insert(0, mg.REF)
  from ffn.inference import inference_utils as ref_utils
  from ffn.inference import seed as ref_seed
  import types
  from scipy import ndimage
  # un-vendored: skimage.morphology.binary_erosion(image) = erosion by the connectivity-1 cross with everything outside
  # the image counted as foreground (scikit-image: ndi.binary_erosion(image, structure=footprint, border_value=True))
  ref_seed.skimage = types.SimpleNamespace(morphology=types.SimpleNamespace(
      binary_erosion=lambda x: ndimage.binary_erosion(x, border_value=True)))

  class FakeCanvas:
    image = make_image()
    shape = SHAPE
    margin = np.asarray(MARGIN)
    restrictor = None
    segmentation = np.zeros(SHAPE, np.int32)

  out = {'image': FakeCanvas.image, 'margin': np.asarray(MARGIN)}
  for name, cls, kwargs in CASES:
    canvas = FakeCanvas()
    policy = getattr(ref_seed, cls)(canvas, **kwargs)
    coords = np.array([tuple(int(v) for v in c) for c in policy], dtype=np.int64).reshape(-1, 3)
    out['coords_' + name] = coords
    print(name, cls, kwargs, '->', coords.shape[0], 'seeds', coords[:2].tolist())
    assert coords.shape[0] > 0

  c = ref_utils.Counters()
  c['inference-calls'].IncrementBy(144)
  c['voxels-segmented'].Set(123456)
  c['inference-time-ms'].IncrementBy(2500)
  sub = c.get_sub_counters()
  sub['skip_invalid_pos'].IncrementBy(7)
  sub['voxels-segmented'].IncrementBy(44)
  out['counters_dumps'] = np.asarray(c.dumps())
  out['sub_counters_dumps'] = np.asarray(sub.dumps())
  d = ref_utils.Counters()
  d.loads(c.dumps())
  out['counters_after_loads'] = np.asarray(d.dumps())
  out['counters_items'] = np.asarray(sorted('%s=%r' % (k, v.value) for k, v in c))
  print(c.dumps())
  np.savez_compressed(os.path.join(HERE, 'seed_policies_ref.npz'), **out)
  print('wrote seed_policies_ref.npz')


if __name__ == '__main__':
  main()
