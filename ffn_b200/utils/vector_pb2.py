"""Drop-in for ffn/utils/vector_pb2.py."""
from ..inference.protos import Vector3d, Vector3f, Vector3j  # noqa: F401
