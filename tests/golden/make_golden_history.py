"""Adds the keep_history fixture: Canvas.history / Canvas.history_deleted of the REAL reference
`Canvas.segment_at` (ffn/inference/inference.py:420-422, :520-521) for the object of segment_at_64.npz.

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_golden_history.py

Same harness as make_golden.py (stubbed third-party imports, oracle network as the executor client);
kept separate so that the existing fixtures are not regenerated.  Output: segment_at_history_64.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
  os.environ.setdefault('PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION', 'python')
  mg.install_stubs()
  sys.path.insert(0, mg.REF)
  from ffn.inference import inference as ref_inference
  from ffn.inference import inference_pb2 as ref_pb2
  from ffn.inference import movement as ref_movement
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

  g = np.load(os.path.join(HERE, 'flood_fill_64.npz'))
  gat = np.load(os.path.join(HERE, 'segment_at_64.npz'))
  image = (g['volume'].astype(np.float32) - 128.0) / 33.0
  opts = ref_pb2.InferenceOptions()
  opts.init_activation = 0.95
  opts.pad_value = 0.05
  opts.move_threshold = 0.9
  opts.segment_threshold = 0.6
  opts.min_segment_size = 1000
  opts.min_boundary_dist.x = opts.min_boundary_dist.y = opts.min_boundary_dist.z = 1
  request = ref_pb2.InferenceRequest()
  request.inference_options.CopyFrom(opts)
  info = ref_model.ModelInfo(np.array([8, 8, 8]), np.array([33, 33, 33]),
                             np.array([33, 33, 33]), np.array([33, 33, 33]))
  canvas = ref_inference.Canvas(info, Client(), image, opts,
                                movement_policy_fn=ref_movement.get_policy_fn(request, info), keep_history=True)
  start = tuple(int(v) for v in gat['start'])
  iters = canvas.segment_at(start)
  history = np.asarray(canvas.history, dtype=np.int32).reshape(-1, 3)
  deleted = np.asarray(canvas.history_deleted, dtype=np.int64)
  assert iters == int(gat['iters']) and np.array_equal(history, gat['trace']), 'does not reproduce segment_at_64.npz'
  # a second object started inside the first one: exercises non-zero history_deleted values
  second = tuple(int(v) for v in history[len(history) // 2])
  canvas.segment_at(second)
  np.savez_compressed(os.path.join(HERE, 'segment_at_history_64.npz'), start=np.asarray(start), history=history,
                      history_deleted=deleted, second_start=np.asarray(second),
                      second_history=np.asarray(canvas.history, dtype=np.int32).reshape(-1, 3),
                      second_history_deleted=np.asarray(canvas.history_deleted, dtype=np.int64))
  print('history: %d steps, deleted sum %d max %d; second object %d steps, deleted sum %d' %
        (len(history), int(deleted.sum()), int(deleted.max()), len(canvas.history),
         int(np.sum(canvas.history_deleted))))


if __name__ == '__main__':
  main()
