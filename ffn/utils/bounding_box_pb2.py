"""Alias of ffn_b200.utils.bounding_box_pb2."""
import sys as _sys
from ffn_b200.utils import bounding_box_pb2 as _impl
_sys.modules[__name__] = _impl
