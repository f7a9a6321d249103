"""Alias of ffn_b200.inference.seed."""
import sys as _sys
from ffn_b200.inference import seed as _impl
_sys.modules[__name__] = _impl
