"""Alias of ffn_b200.inference.align."""
import sys as _sys
from ffn_b200.inference import align as _impl
_sys.modules[__name__] = _impl
