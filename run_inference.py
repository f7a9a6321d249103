"""Runs FFN inference within a dense bounding box (single process, one GPU).

Drop-in for the reference entry point (run_inference.py:40-56):

  python run_inference.py \
    --inference_request="$(cat configs/inference_training_sample2.pbtxt)" \
    --bounding_box 'start { x:0 y:0 z:0 } size { x:250 y:250 z:250 }'

writes <segmentation_output_dir>/<x>/<y>/seg-<x>_<y>_<z>.npz (+ .prob) and counters.txt.
"""

import os

from absl import app
from absl import flags
from google.protobuf import text_format

from ffn.inference import inference_flags
from ffn.inference import runner as runner_mod
from ffn.utils import bounding_box_pb2

FLAGS = flags.FLAGS

flags.DEFINE_string('bounding_box', None, 'BoundingBox proto in text format defining the area to segmented.')
flags.DEFINE_integer('device', 0, 'CUDA device ordinal.')


def main(unused_argv):
  request = inference_flags.request_from_flags()
  os.makedirs(request.segmentation_output_dir, exist_ok=True)

  bbox = bounding_box_pb2.BoundingBox()
  text_format.Parse(FLAGS.bounding_box, bbox)

  runner = runner_mod.Runner(device=FLAGS.device)
  runner.start(request)
  runner.run((bbox.start.z, bbox.start.y, bbox.start.x), (bbox.size.z, bbox.size.y, bbox.size.x))

  counter_path = os.path.join(request.segmentation_output_dir, 'counters.txt')
  if not os.path.exists(counter_path):
    runner.counters.dump(counter_path)


if __name__ == '__main__':
  app.run(main)
