"""Alias of ffn_b200.inference.resegmentation."""
import sys as _sys
from ffn_b200.inference import resegmentation as _impl
_sys.modules[__name__] = _impl
