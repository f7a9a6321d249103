"""Fixture for shift-mask restriction: MovementRestrictor.is_valid_pos of the REAL reference
(ffn/inference/movement.py:247-336) evaluated at every position of small volumes.

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden_restrictor.py

Output: restrictor_shift.npz (cases: shift field, mask, fov box, scale, threshold -> valid grid).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
  os.environ.setdefault('PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION', 'python')
  mg.install_stubs()
  sys.path.insert(0, mg.REF)
  from ffn.inference import movement as ref_movement

  rng = np.random.RandomState(17)
  out = {}
  n = 0
  for scale, shape in ((1, (9, 20, 22)), (2, (7, 21, 19)), (4, (6, 33, 30))):
    sm_shape = (shape[0], -(-shape[1] // scale), -(-shape[2] // scale))
    for start, size in (((-4, -3, -1), (9, 7, 3)), ((-2, -2, 0), (4, 5, 1)), ((-16, -16, -16), (33, 33, 33)),
                        ((-6, 1, -2), (3, 2, 2))):
      shift = (rng.randn(2, *sm_shape) * 1.7).astype(np.float32)
      mask = rng.rand(*shape) > 0.92
      threshold = 4
      fov = types.SimpleNamespace(start=np.array(start), end=np.array(start) + np.array(size))   # xyz
      restrictor = ref_movement.MovementRestrictor(mask=mask, shift_mask=shift, shift_mask_fov=fov,
                                                   shift_mask_threshold=threshold, shift_mask_scale=scale)
      valid = np.zeros(shape, dtype=bool)
      for z in range(shape[0]):
        for y in range(shape[1]):
          for x in range(shape[2]):
            valid[z, y, x] = restrictor.is_valid_pos((z, y, x))
      out.update({'shift_%d' % n: shift, 'mask_%d' % n: mask, 'fov_start_%d' % n: np.array(start),
                  'fov_size_%d' % n: np.array(size), 'scale_%d' % n: scale, 'threshold_%d' % n: threshold,
                  'valid_%d' % n: valid})
      print('case %d scale %d fov %s+%s: %.1f%% valid' % (n, scale, start, size, 100 * valid.mean()))
      n += 1
  np.savez_compressed(os.path.join(HERE, 'restrictor_shift.npz'), n=n, **out)


if __name__ == '__main__':
  main()
