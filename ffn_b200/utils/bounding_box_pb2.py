"""Drop-in for ffn/utils/bounding_box_pb2.py."""
from ..inference.protos import BoundingBox, BoundingBoxes  # noqa: F401
