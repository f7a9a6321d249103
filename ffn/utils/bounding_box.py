"""Alias of ffn_b200.utils.bounding_box."""
import sys as _sys
from ffn_b200.utils import bounding_box as _impl
_sys.modules[__name__] = _impl
