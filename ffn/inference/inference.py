"""Alias of ffn_b200.inference.inference."""
import sys as _sys
from ffn_b200.inference import inference as _impl
_sys.modules[__name__] = _impl
