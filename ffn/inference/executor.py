"""Alias of ffn_b200.inference.executor."""
import sys as _sys
from ffn_b200.inference import executor as _impl
_sys.modules[__name__] = _impl
