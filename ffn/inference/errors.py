"""Alias of ffn_b200.inference.errors."""
import sys as _sys
from ffn_b200.inference import errors as _impl
_sys.modules[__name__] = _impl
