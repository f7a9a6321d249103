"""`storage.build_mask` fixture from the REAL reference (ffn/inference/storage.py:323-411, unmodified; harness as in
make_golden.py): MaskConfig lists covering the three sources — coordinate expression, image channels (min / max,
invert), volume channels (value lists on a 4-d volume, per-channel and per-config invert) — and their OR.

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden_build_mask.py  ->  build_mask_ref.npz

The mask volume is handed over through `mask_volume_map` (keyed by the serialized DecoratedVolume, exactly the cache
the reference keeps), so no volume backend is needed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

CORNER = (3, 4, 5)
SIZE = (10, 12, 14)
CASES = {
    'expr': ['coordinate_expression { expression: "(x + y > 24) & (z % 3 == 0)" }'],
    'image': ['image { channels { channel: 0 min_value: 0.25 max_value: 0.6 } }'],
    'image_inv': ['image { channels { channel: 0 min_value: -1 max_value: 0.5 invert: true } } invert: true'],
    'volume': ['volume { mask { hdf5: "m.h5:mask" } channels { channel: 1 values: 2 values: 5 } '
               'channels { channel: 0 min_value: 7 max_value: 9 invert: true } }'],
    'all': ['coordinate_expression { expression: "x < 8" }',
            'image { channels { channel: 0 min_value: 0.7 max_value: 10 } }',
            'volume { mask { hdf5: "m.h5:mask" } channels { channel: 1 values: 3 } } invert: true'],
}


def arrays():
  rng = np.random.RandomState(21)
  volume = rng.randint(0, 10, size=(2, 20, 24, 28)).astype(np.uint8)      # 4-d mask volume (channel, z, y, x)
  image = rng.rand(*SIZE).astype(np.float32)
  return volume, image


def main():
  os.environ.setdefault('PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION', 'python')
  mg.install_stubs()
  sys.path.insert(0, mg.REF)
  from google.protobuf import text_format
  from ffn.inference import inference_pb2 as ref_pb2
  from ffn.inference import storage as ref_storage
  from ffn.utils import bounding_box as ref_bbox
  # storage.py imports the un-vendored connectomics.common.bounding_box (stubbed here); the reference's own vendored
  # ffn/utils/bounding_box.py — the class the connectomics one was split out of — provides BoundingBox / intersection
  class _BBox(ref_bbox.BoundingBox):
    def intersection(self, other):          # connectomics spells the vendored module-level intersection() as a method
      return ref_bbox.intersection(self, other)
  import types
  ref_storage.bounding_box = types.SimpleNamespace(BoundingBox=_BBox)
  volume, image = arrays()
  out = {'volume': volume, 'image': image, 'corner': np.asarray(CORNER), 'size': np.asarray(SIZE)}
  for name, texts in CASES.items():
    configs = [text_format.Parse(t, ref_pb2.MaskConfig()) for t in texts]
    cache = {c.volume.mask.SerializeToString(): volume for c in configs if c.WhichOneof('source') == 'volume'}
    mask = ref_storage.build_mask(configs, CORNER, SIZE, mask_volume_map=cache, image=image)
    out['mask_' + name] = np.asarray(mask, dtype=bool)
    print(name, mask.shape, int(mask.sum()), 'of', mask.size)
  np.savez_compressed(os.path.join(HERE, 'build_mask_ref.npz'), **out)
  print('wrote build_mask_ref.npz')


if __name__ == '__main__':
  main()
