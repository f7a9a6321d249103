"""Alias of ffn_b200.inference.movement."""
import sys as _sys
from ffn_b200.inference import movement as _impl
_sys.modules[__name__] = _impl
