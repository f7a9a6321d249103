"""Alias of ffn_b200.inference.runner."""
import sys as _sys
from ffn_b200.inference import runner as _impl
_sys.modules[__name__] = _impl
