"""Command-line flags carrying text-format protos, and their parsers.

Public surface of ffn/inference/inference_flags.py:38-57: the flags `--inference_request` and
`--inference_options`, and `request_from_flags()` / `options_from_flags()` returning the parsed message
(an empty message when the flag is unset).
"""

from absl import flags
from google.protobuf import text_format

from . import inference_pb2

FLAGS = flags.FLAGS

# flag name -> message type it is parsed into
_PROTO_FLAGS = {
    'inference_request': inference_pb2.InferenceRequest,
    'inference_options': inference_pb2.InferenceOptions,
}
for _name, _cls in _PROTO_FLAGS.items():
  flags.DEFINE_string(_name, None, '%s proto in text format.' % _cls.DESCRIPTOR.name)


def _parse_flag(name):
  message = _PROTO_FLAGS[name]()
  text = getattr(FLAGS, name)
  if text:
    text_format.Parse(text, message)
  return message


def options_from_flags():
  return _parse_flag('inference_options')


def request_from_flags():
  return _parse_flag('inference_request')
