"""Alias of ffn_b200.inference.inference_flags."""
import sys as _sys
from ffn_b200.inference import inference_flags as _impl
_sys.modules[__name__] = _impl
