"""Alias of ffn_b200.inference.storage."""
import sys as _sys
from ffn_b200.inference import storage as _impl
_sys.modules[__name__] = _impl
